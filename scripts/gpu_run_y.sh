#!/bin/bash
# GPU call Y (1 GPU): whole-row attention forward — compile-time variants (1: K/V shared by a head's query tiles, 2: deferred epilogue,
# 4: TMA store, 8: two-pass softmax re-reading TMEM) against the per-tile kernel, same box; parity tests for two variants
mkdir -p gpurun_out
for f in 0 2 4 6 7 8 9 10 12 14 15; do
  echo -n "flags=$f  "; DLB_ATTN_ROW_FLAGS=$f timeout 100 python scripts/bench_attn.py 2>&1 | head -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['dlb_attn_fwd_tc_ms'], d['dlb_attn_fwd_tc_tflops'])"
done
echo -n "per-tile  "; DLB_ATTN_NO_ROW=1 timeout 100 python scripts/bench_attn.py 2>&1 | head -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['dlb_attn_fwd_tc_ms'], d['dlb_attn_fwd_tc_tflops'])"
for f in 14 15; do
  echo "== tests flags=$f"; DLB_ATTN_ROW_FLAGS=$f timeout 200 python -m pytest tests/test_attention_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1
done
