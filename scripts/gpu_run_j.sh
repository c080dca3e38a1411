#!/bin/bash
# GPU call J: ncu of the lean elementwise kernels; the other BASELINE configs through bench.py --config
mkdir -p gpurun_out
echo "== ncu: elementwise kernels (lean)"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'ln_modulate|qknorm|gate_residual' -c 14 -o gpurun_out/ncu_elem_lean_r2 -f python scripts/profile_step.py --depth 1 > gpurun_out/ncu_elem.log 2>&1
echo "rc=$? $(ls -la gpurun_out/ncu_elem_lean_r2.ncu-rep 2>/dev/null)"
for cfg in cifar10 txt_to_img sprint; do
  echo "== bench --config $cfg"
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/bench_r2_${cfg}_1gpu.json 2> gpurun_out/bench_r2_${cfg}_1gpu.err
  echo "rc=$?"; tail -n 3 gpurun_out/bench_r2_${cfg}_1gpu.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_${cfg}_1gpu.json').read().strip().splitlines()[-1])
print(d['metric'], d['value'], d['ms_per_step'], 'e2e', d['e2e'], 'mfu', d['roofline']['step_model_flops_frac'], d['loss_check'])
print({k:v for k,v in (d['cpu_baseline'] or {}).items() if k not in ('sample','ref_gpu_what')})
PY
done
echo "== bench --config sprint at a saturating batch (256 / GPU)"
timeout 900 python bench.py --config sprint --batch 256 --steps 10 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/bench_r2_sprint_b256_1gpu.json 2> gpurun_out/bench_r2_sprint_b256.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_sprint_b256_1gpu.json').read().strip().splitlines()[-1])
print(d['metric'], d['value'], d['ms_per_step'], 'mfu', d['roofline']['step_model_flops_frac'])
PY
du -sh gpurun_out
