#!/bin/bash
# The driver's round-end launch line at N = 8 with DEFAULT flags (sampling sweep included) on the final tree
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 6 --warmup 3 \
  > gpurun_out/bench_r2_8gpu_default.json 2> gpurun_out/bench_r2_8gpu_default.err
echo "rc=$?"; tail -n 3 gpurun_out/bench_r2_8gpu_default.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_8gpu_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e'], d['config'].get('dp'), d['gpu_launches'], d['clocks'])
PY
