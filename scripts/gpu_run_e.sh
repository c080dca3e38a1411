#!/bin/bash
# GPU call E: lean elementwise kernels: parity (kernel + model level), microbench, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_e.log 2>&1
echo "rc=$? $(tail -n 3 gpurun_out/pytest_e.log | tr '\n' ' ')"
grep -E "FAILED|Error|assert" gpurun_out/pytest_e.log | cut -c1-300 | head -20
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_configs_gpu.py tests/test_training_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_e2.log 2>&1
echo "rc=$? $(tail -n 3 gpurun_out/pytest_e2.log | tr '\n' ' ')"
grep -E "FAILED|Error|assert" gpurun_out/pytest_e2.log | cut -c1-300 | head -20
echo "== elementwise microbench (lean)"
timeout 300 python scripts/bench_elementwise.py > gpurun_out/elementwise_microbench_r2.jsonl 2> gpurun_out/elementwise.err
cat gpurun_out/elementwise_microbench_r2.jsonl; tail -n 3 gpurun_out/elementwise.err
echo "== elementwise microbench (general kernels, DLB_NO_LEAN=1)"
DLB_NO_LEAN=1 timeout 300 python scripts/bench_elementwise.py > gpurun_out/elementwise_microbench_r2_nolean.jsonl 2>> gpurun_out/elementwise.err
cat gpurun_out/elementwise_microbench_r2_nolean.jsonl
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sample > gpurun_out/bench_r2_e_1gpu.json 2> gpurun_out/bench_r2_e_1gpu.err
echo "rc=$? $(cut -c1-200 gpurun_out/bench_r2_e_1gpu.json)"; tail -n 3 gpurun_out/bench_r2_e_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_e_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['ms_per_step_by_family'])
PY
