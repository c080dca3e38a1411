#!/bin/bash
# GPU call C (round 2): swizzled attention path: probe, parity tests, microbench; small ncu captures (< 64 MiB in total).
mkdir -p gpurun_out
echo "== sw probe + attention parity"
timeout 900 python -m pytest tests/test_attn_sw_probe_gpu.py tests/test_attention_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_c.log 2>&1
echo "rc=$? $(tail -n 3 gpurun_out/pytest_c.log | tr '\n' ' ')"
grep -E "FAILED|Error|assert" gpurun_out/pytest_c.log | cut -c1-300 | head -20
echo "== attention microbench"
timeout 300 python scripts/bench_attn.py > gpurun_out/attn_microbench_r2.jsonl 2> gpurun_out/attn_microbench.err
cat gpurun_out/attn_microbench_r2.jsonl; tail -n 3 gpurun_out/attn_microbench.err
echo "== model-level tests"
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_configs_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_c2.log 2>&1
echo "rc=$? $(tail -n 3 gpurun_out/pytest_c2.log | tr '\n' ' ')"
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_c_1gpu.json 2> gpurun_out/bench_r2_c_1gpu.err
echo "rc=$? $(cut -c1-300 gpurun_out/bench_r2_c_1gpu.json)"; tail -n 3 gpurun_out/bench_r2_c_1gpu.err
echo "== ncu: attention kernels with source"
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'attn_' -c 3 -o gpurun_out/ncu_attn_r2 -f python scripts/profile_step.py --depth 1 > gpurun_out/ncu_attn.log 2>&1
echo "rc=$? $(ls -la gpurun_out/ncu_attn_r2.ncu-rep 2>/dev/null)"
echo "== ncu: elementwise kernels"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'ln_modulate|qknorm|gate_residual|adamw' -c 14 -o gpurun_out/ncu_elem_r2 -f python scripts/profile_step.py --depth 1 > gpurun_out/ncu_elem.log 2>&1
echo "rc=$? $(ls -la gpurun_out/ncu_elem_r2.ncu-rep 2>/dev/null)"
echo "== ncu: light list of a depth-4 step"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --csv --log-file gpurun_out/ncu_step_light_r2.csv python scripts/profile_step.py --depth 4 > gpurun_out/ncu_light.log 2>&1
echo "rc=$? lines=$(wc -l < gpurun_out/ncu_step_light_r2.csv)"
du -sh gpurun_out
