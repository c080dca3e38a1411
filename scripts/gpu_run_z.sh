#!/bin/bash
# GPU call Z (1 GPU): final evidence of round 2 — full GPU test tier, bench (default flags), microbenchmarks, ncu launch list of one
# full-depth step, ncu --set full summaries (GEMM per shape, attention, row kernels, AdamW). Reports are summarised on the box.
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1
echo "rc=$? $(tail -n 3 gpurun_out/pytest_gpu_final.log | tr '\n' ' ')"
grep -E "FAILED|Error|assert" gpurun_out/pytest_gpu_final.log | cut -c1-300 | head -20
echo "== ncu --set full, one launch per GEMM shape"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:gemm -o gpurun_out/ncu_gemm_shapes_r2 -f \
  python scripts/ncu_gemm_shapes.py gpurun_out/ncu_gemm_shapes_manifest.json > gpurun_out/ncu_gemm_shapes.log 2>&1
echo "rc=$?"; tail -n 2 gpurun_out/ncu_gemm_shapes.log
python scripts/ncu_gemm_traffic.py gpurun_out/ncu_gemm_shapes_r2.ncu-rep gpurun_out/ncu_gemm_shapes_manifest.json > gpurun_out/ncu_gemm_traffic_r2.json 2> gpurun_out/ncu_gemm_traffic.err
echo "rc=$?"; head -c 600 gpurun_out/ncu_gemm_traffic_r2.json; tail -n 3 gpurun_out/ncu_gemm_traffic.err
python scripts/ncu_summary.py gpurun_out/ncu_gemm_shapes_r2.ncu-rep > gpurun_out/ncu_gemm_shapes_r2.txt 2>/dev/null
mkdir -p profiles; cp gpurun_out/ncu_gemm_traffic_r2.json profiles/ncu_gemm_traffic_r2.json   # bench.py reads it for roofline.traffic
rm -f gpurun_out/ncu_gemm_shapes_r2.ncu-rep
echo "== bench (default flags)"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final_1gpu.json 2> gpurun_out/bench_r2_final_1gpu.err
echo "rc=$?"; tail -n 3 gpurun_out/bench_r2_final_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e'], d['clocks'])
print(d['roofline']['ms_per_step_by_family'], 'traffic', d['roofline']['traffic'])
print({k:v for k,v in d['cpu_baseline'].items() if k not in ('sample','ref_gpu_what')}, d['loss_check'])
PY
echo "== microbenchmarks"
timeout 300 python scripts/bench_elementwise.py > gpurun_out/elementwise_microbench_r2.jsonl 2>/dev/null
DLB_NO_LEAN=1 timeout 300 python scripts/bench_elementwise.py > gpurun_out/elementwise_microbench_r2_general_kernels.jsonl 2>/dev/null
paste -d' ' <(python -c "
import json
for l in open('gpurun_out/elementwise_microbench_r2.jsonl'):
    d=json.loads(l); print(d['kernel'], d['ms'], d['frac_of_measured_hbm_peak'])") <(python -c "
import json
for l in open('gpurun_out/elementwise_microbench_r2_general_kernels.jsonl'):
    d=json.loads(l); print('| general:', d['ms'], d['frac_of_measured_hbm_peak'])")
timeout 300 python scripts/bench_gemm_pair.py > gpurun_out/gemm_pair_microbench_r2.jsonl 2>/dev/null; cut -c1-600 gpurun_out/gemm_pair_microbench_r2.jsonl
timeout 300 python scripts/bench_attn.py > gpurun_out/attn_microbench_r2.jsonl 2>/dev/null; cut -c1-400 gpurun_out/attn_microbench_r2.jsonl | head -5
echo "== ncu launch list, one full-depth step"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_launches_r2.csv \
  python scripts/profile_step.py --depth 28 > gpurun_out/ncu_list.log 2>&1
echo "rc=$? lines=$(wc -l < gpurun_out/ncu_launches_r2.csv)"
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/ncu_launches_r2.csv', errors='ignore')) if len(r)>5]
hdr=next(r for r in rows if 'Kernel Name' in r); i0=rows.index(hdr)
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[i0+1:]:
    try: v=float(r[mv].replace(',',''))
    except Exception: continue
    v*= {'ns':1e-3,'us':1.0,'ms':1e3}.get(r[mu],1.0)
    name=re.sub(r'<.*','',r[kn]).split('(')[0].strip()
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
with open('gpurun_out/ncu_launch_summary_r2.txt','w') as f:
    f.write(f"# one DiT-XL/2 + REPA train step (B 128), ncu --metrics gpu__time_duration.sum --clock-control none: {sum(v[0] for v in agg.values())} launches, {tot/1e3:.1f} ms of kernel time (cold-cache, serialised)\n")
    for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        f.write(f"{t/1e3:9.3f} ms {100*t/tot:6.2f} % {n:6d} launches  {k}\n")
print(open('gpurun_out/ncu_launch_summary_r2.txt').read()[:1800])
PY
echo "== ncu --set full: attention, row kernels, AdamW (depth 1)"
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none \
  -k regex:'attn_|ln_modulate|qknorm|gate_residual|adamw' -c 26 -o gpurun_out/ncu_hot_r2 -f python scripts/profile_step.py --depth 1 > gpurun_out/ncu_hot.log 2>&1
echo "rc=$? $(ls -la gpurun_out/ncu_hot_r2.ncu-rep 2>/dev/null)"
python scripts/ncu_summary.py gpurun_out/ncu_hot_r2.ncu-rep > gpurun_out/ncu_hot_kernels_r2.txt 2>/dev/null
python scripts/ncu_sass_stalls.py gpurun_out/ncu_hot_r2.ncu-rep attn_fwd_ws > gpurun_out/ncu_attention_stalls_r2.txt 2>/dev/null
rm -f gpurun_out/ncu_hot_r2.ncu-rep
du -sh gpurun_out
echo "== reference-GPU arm with torch.compile (oracle port under bf16 autocast, compiled blocks)"
timeout 1200 python bench.py --steps 4 --warmup 3 --no-sample --ref-gpu compile > gpurun_out/bench_r2_refgpu_compile.json 2> gpurun_out/bench_r2_refgpu_compile.err
echo "rc=$?"; tail -n 2 gpurun_out/bench_r2_refgpu_compile.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r2_refgpu_compile.json').read().strip().splitlines()[-1])
    print({k:v for k,v in d['cpu_baseline'].items() if k.startswith('ref_gpu') and k!='ref_gpu_what'})
except Exception as e:
    print('compile arm failed', e)
PY
du -sh gpurun_out
