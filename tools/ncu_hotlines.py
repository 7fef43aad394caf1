"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: python tools/ncu_hotlines.py dump.csv [topN]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; H = None; func = None
agg = defaultdict(lambda: [0.0, 0.0, ""])   # (file,line) -> [instr, samples, text]
first_func = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1]; continue
    if len(r) >= 2 and r[0] == "Function Name":
        func = r[1]
        if first_func is None: first_func = func
        continue
    if len(r) > 8 and r[0] == "Line No": H = r; continue
    if H is None or len(r) != len(H) or func != first_func: continue
    ie = H.index("Instructions Executed"); ss = H.index("# Samples")
    key = (cur_file.split("/")[-1], r[0])
    try:
        agg[key][0] += float(r[ie] or 0); agg[key][1] += float(r[ss] or 0)
    except ValueError:
        continue
    if r[1].strip(): agg[key][2] = r[1].strip()
print("function:", first_func[:90])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp-instructions {ti:.4g}, samples {ts:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*v[0]/ti:6.2f}% inst {100*v[1]/ts:6.2f}% smp | {k[0]}:{k[1]:>4} | {v[2][:105]}")
