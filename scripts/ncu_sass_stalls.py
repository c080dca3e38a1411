"""Aggregate `ncu --page source --print-source sass --csv` output: per kernel, share of warp-stall samples by reason and
the SASS instructions that collect the most samples.   usage: ncu -i X.ncu-rep --page source --csv --print-source sass | python scripts/ncu_sass_stalls.py [topN]"""
import csv
import sys

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
secs, cur = [], None
for r in csv.reader(sys.stdin):
    if r and r[0] == "Kernel Name":
        cur = {"fn": r[1], "rows": []}
        secs.append(cur)
    elif r and r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and r and "hdr" in cur:
        cur["rows"].append(r)
for s in secs:
    h = s["hdr"]
    iS, iSrc = h.index("# Samples"), h.index("Source")
    stall_idx = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[iS]) for r in s["rows"] if r[iS].isdigit())
    print("==", s["fn"][:70], "samples", tot, "instructions", len(s["rows"]))
    agg = {}
    for r in s["rows"]:
        for i, n in stall_idx:
            if r[i].isdigit():
                agg[n] = agg.get(n, 0) + int(r[i])
    print("  ", {k: round(v / max(tot, 1), 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(s["rows"], key=lambda r: -int(r[iS]) if r[iS].isdigit() else 0)[:top_n]:
        st = sorted([(int(r[i]), n) for i, n in stall_idx if r[i].isdigit() and int(r[i]) > 0], reverse=True)[:2]
        print("  ", r[iS].rjust(6), f"{int(r[iS]) / max(tot, 1):.3f}", r[iSrc].strip()[:80].ljust(80), st)
