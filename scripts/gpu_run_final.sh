#!/bin/bash
# Final certification of the tree (1 GPU): full GPU test tier, smoke(), default bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1
echo "rc=$? $(tail -n 2 gpurun_out/pytest_gpu_final.log | tr '\n' ' ')"
grep -E "FAILED|Error" gpurun_out/pytest_gpu_final.log | cut -c1-300 | head
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final_1gpu.json 2> gpurun_out/bench_r2_final_1gpu.err
echo "rc=$?"; tail -n 2 gpurun_out/bench_r2_final_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e'], d['clocks'], 'launches', d['gpu_launches'])
print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['step_model_flops_frac'], d['roofline']['ms_per_step_by_family'])
print({k:v for k,v in d['cpu_baseline'].items() if k not in ('sample','ref_gpu_what')}, d['loss_check'])
PY
