#!/bin/bash
# Run each GPU test function in its own process so that one trapped kernel (dead CUDA context)
# does not mask the results of the others. Logs land in gpurun_out/.
# usage: scripts/gpu_pytest_isolated.sh tests/test_gemm_gpu.py [more files]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/nvsmi.txt 2>&1
for f in "$@"; do
  for t in $(python -m pytest "$f" -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u); do
    name=$(echo "$t" | tr '/:' '__')
    timeout 300 python -m pytest "$t" -m gpu -q --timeout 120 -p no:cacheprovider > "gpurun_out/${name}.log" 2>&1
    echo "$t rc=$? $(tail -n 1 gpurun_out/${name}.log)"
  done
done
