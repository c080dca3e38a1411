#!/bin/bash
# ncu --set full of the whole-row attention forward kernel inside a depth-1 train step (summary + warp-stall samples per SASS line)
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'attn_fwd_row' -c 2 -o gpurun_out/ncu_attn_row_r2 -f \
  python scripts/profile_step.py --depth 1 > gpurun_out/ncu_attn_row.log 2>&1
echo "rc=$? $(ls -la gpurun_out/ncu_attn_row_r2.ncu-rep 2>/dev/null)"
python scripts/ncu_summary.py gpurun_out/ncu_attn_row_r2.ncu-rep > gpurun_out/ncu_attn_fwd_row_r2.txt 2>/dev/null
ncu -i gpurun_out/ncu_attn_row_r2.ncu-rep --page source --csv --print-source sass 2>/dev/null | python scripts/ncu_sass_stalls.py 16 > gpurun_out/ncu_attn_fwd_row_stalls_r2.txt 2>/dev/null
head -30 gpurun_out/ncu_attn_fwd_row_r2.txt; head -30 gpurun_out/ncu_attn_fwd_row_stalls_r2.txt | cut -c1-200
rm -f gpurun_out/ncu_attn_row_r2.ncu-rep
