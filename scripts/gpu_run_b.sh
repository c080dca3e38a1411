#!/bin/bash
# GPU call B (round 2): swizzled-tile probe + TMA stream rates, optimizer tests, ncu captures incl. the backward kernels.
mkdir -p gpurun_out
echo "== sw probe"
timeout 600 python -m pytest tests/test_attn_sw_probe_gpu.py tests/test_umma_probe_gpu.py tests/test_training_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_b.log 2>&1
echo "rc=$? $(tail -n 3 gpurun_out/pytest_b.log | tr '\n' ' ')"
grep -E "FAILED|Error|assert" gpurun_out/pytest_b.log | cut -c1-300 | head -30
echo "== tma stream rates"
timeout 300 python scripts/bench_sw_stream.py > gpurun_out/tma_stream_r2.jsonl 2> gpurun_out/tma_stream.err
cat gpurun_out/tma_stream_r2.jsonl; tail -n 3 gpurun_out/tma_stream.err
echo "== ncu launch list of one profiled step (depth 4)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --csv --log-file gpurun_out/ncu_step_light_r2.csv python scripts/profile_step.py --depth 4 > gpurun_out/ncu_light.log 2>&1
echo "rc=$? lines=$(wc -l < gpurun_out/ncu_step_light_r2.csv)"
echo "== ncu --set full on the non-GEMM hot kernels (depth 1)"
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none \
  -k regex:'attn_|ln_modulate|qknorm|gate_residual|adamw' -c 24 -o gpurun_out/ncu_hot_r2 -f python scripts/profile_step.py --depth 1 > gpurun_out/ncu_hot.log 2>&1
echo "rc=$? $(ls -la gpurun_out/ncu_hot_r2.ncu-rep 2>/dev/null)"
timeout 600 ncu --profile-from-start off --set full --clock-control none \
  -k regex:'gemm' -c 40 -o gpurun_out/ncu_gemm_r2 -f python scripts/profile_step.py --depth 1 > gpurun_out/ncu_gemm.log 2>&1
echo "rc=$? $(ls -la gpurun_out/ncu_gemm_r2.ncu-rep 2>/dev/null)"
