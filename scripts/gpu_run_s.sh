#!/bin/bash
# GPU call S (8 GPUs): copy-engine gradient reducer with per-bucket CUDA graphs at N = 8 against the same-box 1-GPU step
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8gpu.txt 2>&1
runN() {  # n, tag, extra args
  n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 8 --warmup 3 \
    --no-sample --no-cpu-baseline "$@" > gpurun_out/dp${n}_${tag}.json 2> gpurun_out/dp${n}_${tag}.err
  rc=$?
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/dp${n}_${tag}.json').read().strip().splitlines()[-1])
    dp=d['config'].get('dp',{})
    f=d['roofline']['ms_per_step_by_family']
    print('N=${n} ${tag}', 'rc=$rc', 'img/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gemm TF/s', d['roofline']['achieved'], 'mode', dp.get('mode'), 'exposed', dp.get('exposed_tail_ms'), 'buckets', dp.get('buckets'), 'clk', d['clocks'].get('sm_mhz'))
except Exception as e:
    print('N=${n} ${tag}', 'rc=$rc', 'FAILED', e)
    import subprocess; print(subprocess.run('tail -n 15 gpurun_out/dp${n}_${tag}.err', shell=True, capture_output=True, text=True).stdout)
PY
  cp gpurun_out/dp_timeline_${n}gpu.json gpurun_out/dp_timeline_${n}gpu_${tag}.json 2>/dev/null
}
echo "== 1 GPU on this box"
timeout 600 python bench.py --steps 8 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/dp1_samebox_s.json 2> gpurun_out/dp1_samebox_s.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/dp1_samebox_s.json').read().strip().splitlines()[-1])
print('1gpu img/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gemm TF/s', d['roofline']['achieved'], 'clk', d['clocks'].get('sm_mhz'))
PY
runN 8 ce_graph --dp-mode ce
runN 8 ce_graph_b256 --dp-mode ce --bucket-mb 256
du -sh gpurun_out
