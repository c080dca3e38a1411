#!/bin/bash
mkdir -p gpurun_out
echo "== in-step-like attention timing: swizzled TMA path"
timeout 300 python scripts/bench_attn_instep.py 2>&1 | tail -n 2 | cut -c1-400
echo "== cp.async path (DLB_ATTN_NO_TMA=1)"
DLB_ATTN_NO_TMA=1 timeout 300 python scripts/bench_attn_instep.py 2>&1 | tail -n 2 | cut -c1-400
echo "== isolated"
timeout 300 python scripts/bench_attn.py dit_xl2 2>&1 | tail -n 1 | cut -c1-400
DLB_ATTN_NO_TMA=1 timeout 300 python scripts/bench_attn.py dit_xl2 2>&1 | tail -n 1 | cut -c1-400
echo "== elementwise microbench"
timeout 300 python scripts/bench_elementwise.py 2>/dev/null | grep -E "gate|ln_mod|qknorm"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "gate" 2>&1 | tail -n 2
