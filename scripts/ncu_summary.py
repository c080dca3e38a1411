"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics the roofline discussion uses.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/ncu_<what>_rNN.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# {rep}: ncu --set full --clock-control none (per launch; cold caches, serialised)")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(f"\n## {name[:160]}")
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"  {w:75s} {r[i]:>16s} {units[i]}")
