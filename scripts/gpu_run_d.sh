#!/bin/bash
# GPU call D: attention pipeline restructure check (parity + microbench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_d.log 2>&1
echo "rc=$? $(tail -n 3 gpurun_out/pytest_d.log | tr '\n' ' ')"
grep -E "FAILED|Error|assert" gpurun_out/pytest_d.log | cut -c1-300 | head -20
timeout 300 python scripts/bench_attn.py > gpurun_out/attn_microbench_r2.jsonl 2> gpurun_out/attn_microbench.err
cat gpurun_out/attn_microbench_r2.jsonl; tail -n 3 gpurun_out/attn_microbench.err
