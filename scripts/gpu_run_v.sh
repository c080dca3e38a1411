#!/bin/bash
# GPU call V (2 GPUs): the driver's round-end launch lines with DEFAULT flags at N = 2 (our arm and the reference arm)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 \
  > gpurun_out/bench_r2_2gpu_default.json 2> gpurun_out/bench_r2_2gpu_default.err
echo "rc=$?"; tail -n 3 gpurun_out/bench_r2_2gpu_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_2gpu_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e'], d['config'].get('dp'), d['gpu_launches'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 2 --steps 1 --warmup 3 \
  > gpurun_out/bench_r2_2gpu_reference_arm.json 2> gpurun_out/bench_r2_2gpu_reference_arm.err
echo "rc=$?"; cut -c1-400 gpurun_out/bench_r2_2gpu_reference_arm.json
for cfg in cifar10 txt_to_img sprint; do
  CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/bench_r2_${cfg}_1gpu.json 2> gpurun_out/bench_r2_${cfg}_1gpu.err &
  CUDA_VISIBLE_DEVICES=1 true
  wait
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_${cfg}_1gpu.json').read().strip().splitlines()[-1])
print('${cfg}', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'mfu', d['roofline']['step_model_flops_frac'])
PY
done
