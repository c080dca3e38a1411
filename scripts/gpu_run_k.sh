#!/bin/bash
# GPU call K (2 GPUs): data-parallel overlap sweep (NCCL CTA limit, SMs reserved for NCCL, tail bucket) against the same-box 1-GPU
# step; 2-GPU parity test; learned-variance sampler tests; txt_to_img / sprint benches.
mkdir -p gpurun_out
run2() {  # tag, extra args
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 \
    --no-sample --no-cpu-baseline "$@" > gpurun_out/dp2_${tag}.json 2> gpurun_out/dp2_${tag}.err
  rc=$?
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/dp2_${tag}.json').read().strip().splitlines()[-1])
    dp=d['config'].get('dp',{})
    print('${tag}', 'rc=$rc', 'img/s', d['value'], 'ms', d['ms_per_step'], 'gemm TF/s', d['roofline']['achieved'], 'exposed', dp.get('exposed_tail_ms'), 'buckets', dp.get('buckets'), 'clk', d['clocks'].get('sm_mhz'))
except Exception as e:
    print('${tag}', 'rc=$rc', 'FAILED', e)
PY
  cp gpurun_out/dp_timeline_2gpu.json gpurun_out/dp_timeline_2gpu_${tag}.json 2>/dev/null
}
echo "== 1 GPU on this box"
timeout 600 python bench.py --steps 8 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/dp1_samebox.json 2> gpurun_out/dp1_samebox.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/dp1_samebox.json').read().strip().splitlines()[-1])
print('1gpu img/s', d['value'], 'ms', d['ms_per_step'], 'gemm TF/s', d['roofline']['achieved'], 'clk', d['clocks'].get('sm_mhz'))
PY
echo "== 2 GPUs"
run2 r1like --comm-ctas 0 --reserve-sms 0 --bucket-mb 256 --tail-bucket-mb 0
run2 c4r0 --comm-ctas 4 --reserve-sms 0
run2 c4r4 --comm-ctas 4 --reserve-sms 4
run2 c2r2 --comm-ctas 2 --reserve-sms 2
run2 c8r8 --comm-ctas 8 --reserve-sms 8
run2 c8r0 --comm-ctas 8 --reserve-sms 0
run2 c0r0tail --comm-ctas 0 --reserve-sms 0
echo "== NCCL_DEBUG=INFO (algorithm / channels of the reduce communicator)"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 \
    --no-sample --no-cpu-baseline > gpurun_out/dp2_nccl_info.log 2>&1
grep -i "nvls\|channels\|Connected\|max_ctas\|maxCTAs\|nChannels" gpurun_out/dp2_nccl_info.log | sort | uniq -c | sort -rn | head -20
echo "== tests: dp parity (2 GPUs), gaussian, flow"
timeout 900 python -m pytest tests/test_dp_gpu.py tests/test_gaussian_gpu.py tests/test_flow_gpu.py -x -q -m gpu 2>&1 | tail -5
for cfg in txt_to_img sprint; do
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
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
