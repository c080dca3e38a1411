#!/bin/bash
# GPU call N/O (1 GPU): row-kernel prefetch rotation, SwiGLU-backward epilogue with register-prefetched H — tests, same-box A/B against the pre-packing build, e2e with prefetcher
mkdir -p gpurun_out
echo "== tests (kernels, training, models)"
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_gpu.py tests/test_models_gpu.py tests/test_training_gpu.py tests/test_perceiver_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_o.log 2>&1
echo "rc=$? $(tail -n 3 gpurun_out/pytest_o.log | tr '\n' ' ')"
grep -E "FAILED|Error|assert" gpurun_out/pytest_o.log | cut -c1-300 | head -20
for v in new prev new; do
  echo "== elementwise microbench: $v"
  if [ $v = prev ]; then export DIFFULAB_B200_LIB=$PWD/scripts/ab/libdiffulab_b200_prev.so; else unset DIFFULAB_B200_LIB; fi
  timeout 300 python scripts/bench_elementwise.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()[:200]); continue
    print(' ', d['kernel'], d['ms'], d['frac_of_measured_hbm_peak'])
" | tee gpurun_out/elementwise_o_${v}.txt
done
for v in prev new prev new; do
  echo "== bench (no sample / cpu): $v"
  if [ $v = prev ]; then export DIFFULAB_B200_LIB=$PWD/scripts/ab/libdiffulab_b200_prev.so; else unset DIFFULAB_B200_LIB; fi
  timeout 600 python bench.py --steps 8 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/bench_o_${v}.json 2> gpurun_out/bench_o_${v}.err
  tail -n 2 gpurun_out/bench_o_${v}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_o_${v}.json').read().strip().splitlines()[-1])
f=d['roofline']['ms_per_step_by_family']
print('${v}', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'clk', d['clocks'].get('sm_mhz'), {k:f[k] for k in ('gemm_swiglu_bwd','ln_modulate_bwd','ln_modulate_fwd','qknorm_rope_bwd','qknorm_rope_fwd','gate_residual_bwd','gate_residual_fwd') if k in f})
PY
done
unset DIFFULAB_B200_LIB
du -sh gpurun_out
