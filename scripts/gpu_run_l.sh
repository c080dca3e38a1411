#!/bin/bash
# GPU call L (2 GPUs): peer-memory gradient reducers (copy-engine and NVLS multicast) — parity test, then the overlap sweep
mkdir -p gpurun_out
echo "== dp parity test (nccl / ce / nvls)"
timeout 600 python -m pytest tests/test_dp_gpu.py -x -q -m gpu 2>&1 | tail -15
run2() {  # tag, extra args
  tag=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 \
    --no-sample --no-cpu-baseline "$@" > gpurun_out/dp2_${tag}.json 2> gpurun_out/dp2_${tag}.err
  rc=$?
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/dp2_${tag}.json').read().strip().splitlines()[-1])
    dp=d['config'].get('dp',{})
    f=d['roofline']['ms_per_step_by_family']
    print('${tag}', 'rc=$rc', 'img/s', d['value'], 'ms', d['ms_per_step'], 'gemm TF/s', d['roofline']['achieved'], 'exposed', dp.get('exposed_tail_ms'), 'buckets', dp.get('buckets'), 'clk', d['clocks'].get('sm_mhz'), 'famsum', round(sum(f.values()),2))
except Exception as e:
    print('${tag}', 'rc=$rc', 'FAILED', e)
    import subprocess; print(subprocess.run('tail -n 12 gpurun_out/dp2_${tag}.err', shell=True, capture_output=True, text=True).stdout)
PY
  cp gpurun_out/dp_timeline_2gpu.json gpurun_out/dp_timeline_2gpu_${tag}.json 2>/dev/null
}
echo "== 1 GPU on this box"
timeout 600 python bench.py --steps 8 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/dp1_samebox_l.json 2> gpurun_out/dp1_samebox_l.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/dp1_samebox_l.json').read().strip().splitlines()[-1])
f=d['roofline']['ms_per_step_by_family']
print('1gpu img/s', d['value'], 'ms', d['ms_per_step'], 'gemm TF/s', d['roofline']['achieved'], 'clk', d['clocks'].get('sm_mhz'), 'famsum', round(sum(f.values()),2), 'swiglu_bwd', f.get('gemm_swiglu_bwd'), 'ln_fwd', f.get('ln_modulate_fwd'))
PY
echo "== 2 GPUs"
run2 nccl_r1like --dp-mode nccl --comm-ctas 0 --reserve-sms 0 --bucket-mb 256 --tail-bucket-mb 0
run2 ce --dp-mode ce --reserve-sms 0
run2 ce_b64 --dp-mode ce --reserve-sms 0 --bucket-mb 64 --tail-bucket-mb 16
run2 nvls_c4 --dp-mode nvls --comm-ctas 4 --reserve-sms 0
run2 nvls_c8 --dp-mode nvls --comm-ctas 8 --reserve-sms 0
run2 nvls_c4r4 --dp-mode nvls --comm-ctas 4 --reserve-sms 4
run2 nvls_c16 --dp-mode nvls --comm-ctas 16 --reserve-sms 0
echo "== gemm + kernels tests (packed fp32 epilogue / LN forward)"
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -4
du -sh gpurun_out
