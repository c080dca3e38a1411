#!/bin/bash
# GPU call T (1 GPU): the other BASELINE configs through bench.py on the final build (prefetcher, final row kernels)
mkdir -p gpurun_out
for cfg in cifar10 txt_to_img sprint; do
  echo "== bench --config $cfg"
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/bench_r2_${cfg}_1gpu.json 2> gpurun_out/bench_r2_${cfg}_1gpu.err
  echo "rc=$?"; tail -n 2 gpurun_out/bench_r2_${cfg}_1gpu.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_${cfg}_1gpu.json').read().strip().splitlines()[-1])
print(d['metric'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'mfu', d['roofline']['step_model_flops_frac'], d['loss_check'])
print({k:v for k,v in (d['cpu_baseline'] or {}).items() if k not in ('sample','ref_gpu_what')})
PY
done
timeout 900 python bench.py --config sprint --batch 256 --steps 10 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/bench_r2_sprint_b256_1gpu.json 2> gpurun_out/bench_r2_sprint_b256.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_sprint_b256_1gpu.json').read().strip().splitlines()[-1])
print(d['metric'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'mfu', d['roofline']['step_model_flops_frac'])
PY
