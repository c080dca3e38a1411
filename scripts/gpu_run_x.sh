#!/bin/bash
# GPU call X (1 GPU): whole-row attention forward kernel — parity tests, then same-box A/B against the per-tile kernel
mkdir -p gpurun_out
echo "== attention tests (row kernel on)"
timeout 300 python -m pytest tests/test_attention_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8
echo "== microbench: row kernel"
timeout 120 python scripts/bench_attn.py 2>&1 | cut -c1-330
echo "== microbench: per-tile kernel (DLB_ATTN_NO_ROW=1)"
DLB_ATTN_NO_ROW=1 timeout 120 python scripts/bench_attn.py 2>&1 | cut -c1-330
