"""ncu report of scripts/ncu_gemm_shapes.py + its manifest -> per-shape DRAM traffic vs algorithmic bytes (bench.py's roofline.traffic).
usage: python scripts/ncu_gemm_traffic.py gpurun_out/ncu_gemm_shapes_r2.ncu-rep gpurun_out/ncu_gemm_shapes_manifest.json > profiles/ncu_gemm_traffic_r2.json"""
import csv
import io
import json
import subprocess
import sys

rep, man = sys.argv[1], json.load(open(sys.argv[2]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return None


units = rows[1]


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


launches = [r for r in rows[2:] if "gemm" in r[col["Kernel Name"]]]
assert len(launches) == len(man), f"{len(launches)} gemm launches in the report, {len(man)} in the manifest"
by = {}
tot = 0.0
for r, m in zip(launches, man):
    rd = to_bytes(num(r, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]])
    wr = to_bytes(num(r, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]])
    dur = num(r, "gpu__time_duration.sum")
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)
    tot += rd + wr
    by[m["label"]] = {"M": m["M"], "N": m["N"], "K": m["K"], "kernel": r[col["Kernel Name"]].split("(")[0][-60:], "grid": r[col["Grid Size"]],
                      "dram_bytes": int(rd + wr), "algorithmic_bytes": int(m["algorithmic_bytes"]), "dram_over_algorithmic": round((rd + wr) / m["algorithmic_bytes"], 3),
                      "duration_us_under_ncu": round(dur_us, 1), "tflops_under_ncu": round(m["flops"] / dur_us / 1e6, 1),
                      "tensor_pipe_pct": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                      "dram_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
print(json.dumps({"source": "ncu --set full --clock-control none, one launch per shape after an L2 flush (scripts/ncu_gemm_shapes.py)",
                  "mean_bytes_per_launch": int(tot / len(man)), "by_shape": by}, indent=1))
