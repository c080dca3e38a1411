#!/bin/bash
# GPU call A (round 2): full GPU test tier, default bench, ncu launch list + per-kernel captures of one profiled step.
# usage (from the repo root, under gpurun): bash scripts/gpu_run_a.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used,power.limit --format=csv > gpurun_out/nvsmi.txt 2>&1
echo "== new tests first (real shapes, training, flow)"
timeout 900 python -m pytest tests/test_configs_gpu.py tests/test_training_gpu.py tests/test_flow_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_new.log 2>&1
echo "rc=$? $(tail -n 3 gpurun_out/pytest_new.log | tr '\n' ' ')"
grep -E "^\[parity|FAILED|Error" gpurun_out/pytest_new.log | cut -c1-600 | head -40
echo "== full gpu tier"
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$? $(tail -n 3 gpurun_out/pytest_gpu.log | tr '\n' ' ')"
echo "== bench (default flags)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_a_1gpu.json 2> gpurun_out/bench_r2_a_1gpu.err
echo "rc=$? $(cut -c1-700 gpurun_out/bench_r2_a_1gpu.json)"
tail -n 5 gpurun_out/bench_r2_a_1gpu.err
echo "== ncu launch list of one profiled step (depth 4)"
timeout 600 ncu --nvtx --nvtx-include "profiled_step/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --csv --log-file gpurun_out/ncu_step_light_r2.csv python scripts/profile_step.py --depth 4 > gpurun_out/ncu_light.log 2>&1
echo "rc=$? lines=$(wc -l < gpurun_out/ncu_step_light_r2.csv)"
echo "== ncu --set full on the non-GEMM hot kernels + a few GEMMs (depth 1)"
timeout 900 ncu --nvtx --nvtx-include "profiled_step/" --set full --import-source on --clock-control none \
  -k regex:'attn_|ln_modulate|qknorm|gate_residual|adamw' -c 24 -o gpurun_out/ncu_hot_r2 -f python scripts/profile_step.py --depth 1 > gpurun_out/ncu_hot.log 2>&1
echo "rc=$? $(ls -la gpurun_out/ncu_hot_r2.ncu-rep 2>/dev/null)"
timeout 600 ncu --nvtx --nvtx-include "profiled_step/" --set full --clock-control none \
  -k regex:'gemm' -c 30 -o gpurun_out/ncu_gemm_r2 -f python scripts/profile_step.py --depth 1 > gpurun_out/ncu_gemm.log 2>&1
echo "rc=$? $(ls -la gpurun_out/ncu_gemm_r2.ncu-rep 2>/dev/null)"
