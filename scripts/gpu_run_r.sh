#!/bin/bash
# GPU call R (2 GPUs): copy-engine reducer with cached peer views + per-bucket CUDA graphs — exact parity test, then bench at N = 2
mkdir -p gpurun_out
echo "== dp parity test (nccl / ce / nvls)"
timeout 600 python -m pytest tests/test_dp_gpu.py -x -q -m gpu 2>&1 | tail -6
run2() {
  tag=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 \
    --no-sample --no-cpu-baseline "$@" > gpurun_out/dp2r_${tag}.json 2> gpurun_out/dp2r_${tag}.err
  rc=$?
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/dp2r_${tag}.json').read().strip().splitlines()[-1])
    dp=d['config'].get('dp',{})
    print('${tag}', 'rc=$rc', 'img/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'mode', dp.get('mode'), 'exposed', dp.get('exposed_tail_ms'), 'clk', d['clocks'].get('sm_mhz'))
except Exception as e:
    print('${tag}', 'rc=$rc', 'FAILED', e)
    import subprocess; print(subprocess.run('tail -n 15 gpurun_out/dp2r_${tag}.err', shell=True, capture_output=True, text=True).stdout)
PY
  grep -i "warn\|capture" gpurun_out/dp2r_${tag}.err | head -3
}
timeout 600 python bench.py --steps 8 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/dp1_samebox_r.json 2> gpurun_out/dp1_samebox_r.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/dp1_samebox_r.json').read().strip().splitlines()[-1])
print('1gpu img/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d['clocks'].get('sm_mhz'))
PY
run2 ce_graph --dp-mode ce
run2 ce_eager --dp-mode ce --no-dp-graphs
