"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line and per phase
(line ranges of tron1_core.cuh).  usage: python tools/ncu_hotlines.py dump.csv [topN]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; H = None; func = None; first_func = None
agg = defaultdict(lambda: defaultdict(float)); text = {}
cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name":
        func = r[1]
        if first_func is None: first_func = func
        continue
    if len(r) > 8 and r[0] == "Line No": H = r; idx = {h: i for i, h in enumerate(H)}; continue
    if H is None or len(r) != len(H) or func != first_func: continue
    if r[0].strip():            # source-line row (aggregated over its SASS)
        cur = (cur_file, int(r[0])); text[cur] = r[1].strip()
        a = agg[cur]
        for name, col in (("inst", "Instructions Executed"), ("smp", "# Samples"), ("noinst", "stall_no_inst"), ("wait", "stall_wait"),
                          ("ssb", "stall_short_sb"), ("lsb", "stall_long_sb"), ("math", "stall_math"), ("sel", "stall_selected")):
            try: a[name] += float(r[idx[col]] or 0)
            except ValueError: pass
    elif cur is not None and r[2].startswith("0x"):
        agg[cur]["sass"] += 1
ti = sum(v["inst"] for v in agg.values()); ts = sum(v["smp"] for v in agg.values()); tsass = sum(v["sass"] for v in agg.values())
print("function:", first_func[:90])
print(f"static SASS {tsass:.0f}, executed warp-instructions {ti:.4g}, samples {ts:.0f}")
PH = [("gait_contact", 57, 75), ("model_step", 133, 168), ("horizon_sums", 173, 204), ("free_response", 205, 231), ("input_response", 232, 283),
      ("adjoint", 284, 333), ("face_basis", 334, 345), ("build_hessian", 346, 434), ("cholesky_regs", 435, 504), ("forward_regs", 505, 536),
      ("backward_regs", 537, 576), ("generic chol/solves", 577, 642), ("project_pyramid", 643, 664), ("gradient", 665, 676),
      ("face_solve", 677, 736), ("check_optimality", 737, 817), ("setup_instance", 818, 852), ("solve_instance(+ADMM)", 853, 985),
      ("make_reference", 986, 1000)]
ph = defaultdict(lambda: defaultdict(float))
for (f, ln), v in agg.items():
    name = f
    if f == "tron1_core.cuh":
        name = next((n for n, a, b in PH if a <= ln <= b), "tron1_core other")
    for k, x in v.items(): ph[name][k] += x
print(f"{'phase':26s} {'SASS':>6s} {'inst%':>6s} {'smp%':>6s} {'noinst':>6s} {'wait':>6s} {'shortsb':>7s} {'longsb':>6s} {'math':>6s}")
for n, v in sorted(ph.items(), key=lambda kv: -kv[1]["smp"]):
    print(f"{n:26s} {v['sass']:6.0f} {100*v['inst']/ti:6.2f} {100*v['smp']/ts:6.2f} {100*v['noinst']/ts:6.2f} {100*v['wait']/ts:6.2f} {100*v['ssb']/ts:7.2f} {100*v['lsb']/ts:6.2f} {100*v['math']/ts:6.2f}")
print()
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    print(f"{v['sass']:5.0f} sass {100*v['inst']/ti:6.2f}% inst {100*v['smp']/ts:6.2f}% smp (noinst {100*v['noinst']/ts:5.2f}) | {k[0]}:{k[1]:>4} | {text[k][:95]}")
