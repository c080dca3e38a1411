#!/bin/bash
# GPU call W2 (1 GPU): gradient memset overlapped with the forward pass (default) vs on the compute stream (DLB_SYNC_ZERO=1), same box
mkdir -p gpurun_out
for v in async sync async sync; do
  if [ $v = sync ]; then export DLB_SYNC_ZERO=1; else unset DLB_SYNC_ZERO; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/bench_w_$v.json 2> gpurun_out/bench_w_$v.err
  tail -n 1 gpurun_out/bench_w_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_w_$v.json').read().strip().splitlines()[-1])
f=d['roofline']['ms_per_step_by_family']
print('$v', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'famsum', round(sum(f.values()),2), 'clk', d['clocks'].get('sm_mhz'))
PY
done
