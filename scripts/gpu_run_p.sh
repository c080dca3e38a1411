#!/bin/bash
# GPU call P (1 GPU): SwiGLU-backward GEMM as a CTA pair (DLB_SWIGLU_BWD_PAIR=1) vs single CTA, same box; gemm tests with the pair path
mkdir -p gpurun_out
echo "== gemm tests with the pair SwiGLU-backward path"
DLB_SWIGLU_BWD_PAIR=1 timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for v in single pair single pair; do
  if [ $v = pair ]; then export DLB_SWIGLU_BWD_PAIR=1; else unset DLB_SWIGLU_BWD_PAIR; fi
  timeout 600 python bench.py --steps 8 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/bench_p_${v}.json 2> gpurun_out/bench_p_${v}.err
  tail -n 2 gpurun_out/bench_p_${v}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_p_${v}.json').read().strip().splitlines()[-1])
f=d['roofline']['ms_per_step_by_family']
print('${v}', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d['clocks'].get('sm_mhz'), {k:f[k] for k in ('gemm_swiglu_bwd','gemm_wgrad','gemm_dgrad','gemm_fwd','gemm_swiglu_fwd','ln_modulate_fwd','qknorm_rope_fwd') if k in f})
PY
done
unset DLB_SWIGLU_BWD_PAIR
du -sh gpurun_out
