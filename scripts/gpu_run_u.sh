#!/bin/bash
# GPU call U (1 GPU): persistent-buffer prefetcher — test, then e2e of the SPRINT config (100 MB of inputs per 22 ms step) and of DiT-XL/2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for cfg in sprint imagenet_repa txt_to_img; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/bench_u_${cfg}.json 2> gpurun_out/bench_u_${cfg}.err
  echo "rc=$?"; tail -n 2 gpurun_out/bench_u_${cfg}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_u_${cfg}.json').read().strip().splitlines()[-1])
print('${cfg}', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'h2d', d['e2e']['h2d_bytes_per_step'])
PY
done
