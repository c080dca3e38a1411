#!/bin/bash
# GPU call I: full GPU tier on the current tree, bench (default flags), SASS-free ncu light list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$? $(tail -n 3 gpurun_out/pytest_gpu.log | tr '\n' ' ')"
grep -E "FAILED|Error|assert" gpurun_out/pytest_gpu.log | cut -c1-300 | head -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_i_1gpu.json 2> gpurun_out/bench_r2_i_1gpu.err
echo "rc=$?"; tail -n 3 gpurun_out/bench_r2_i_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_i_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['ms_per_step_by_family'])
print({k:v for k,v in d['cpu_baseline'].items() if k!='sample' and k!='ref_gpu_what'})
PY
