"""Summarise an ncu report + launch list into profiles/ (tracked evidence).
usage: python tools/ncu_summary.py <tag> [gpurun_out/prof.ncu-rep] [gpurun_out/launches.csv] [--config C --batch B]
With --config the hot kernel's DRAM traffic and executed FP64 work are merged into profiles/traffic.json under
configs[C] (bench.py --config C reads them back as roofline.traffic / roofline.ncu_check)."""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
argv = list(sys.argv)
cfg_id, cfg_batch = None, None
if "--config" in argv:
    i = argv.index("--config"); cfg_id = argv[i + 1]; del argv[i:i + 2]
if "--batch" in argv:
    i = argv.index("--batch"); cfg_batch = int(argv[i + 1]); del argv[i:i + 2]
tag = argv[1]
rep = argv[2] if len(argv) > 2 else os.path.join(ROOT, "gpurun_out", "prof.ncu-rep")
lst = argv[3] if len(argv) > 3 else os.path.join(ROOT, "gpurun_out", "launches.csv")
out = os.path.join(ROOT, "profiles")

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]

summary = {"tag": tag}
if os.path.exists(rep):
    # either an .ncu-rep or the `ncu -i <rep> --page raw --csv` export of one (the reports themselves are too large to
    # bring back from the GPU box: tools/run_profile.sh exports the raw page there and deletes them)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    kcol = H.index("Kernel Name")
    kernels = []
    for r in rows[2:]:
        k = {"kernel": r[kcol]}
        for i, h in enumerate(H):
            if h in KEEP or ("warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")) or "dmma" in h.lower() or \
                    ("pipe_tensor" in h and "pct_of_peak_sustained_active" in h) or "pipe_fp64" in h and "pct" in h:
                k[h] = r[i] + (" " + rows[1][i] if rows[1][i] else "")
        kernels.append(k)
    summary["kernels"] = kernels
    if kernels:
        def num(s):
            return float(s.split()[0].replace(",", ""))
        def dur_us(k):
            v, u = k["gpu__time_duration.sum"].split()[:2]
            return float(v.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        k0 = max(kernels, key=dur_us)   # the hot kernel of the step
        unit_r = k0["dram__bytes_read.sum"].split()[1]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        traffic = num(k0["dram__bytes_read.sum"]) * mult[unit_r] + num(k0["dram__bytes_write.sum"]) * mult[k0["dram__bytes_write.sum"].split()[1]]
        summary["dram_bytes_per_launch"] = traffic
        # executed FP64 work of the hot kernel (first kernel in the report): thread-level DFMA/DMUL/DADD
        try:
            k1 = k0
            cyc = num(k1["sm__cycles_elapsed.max"])
            dfma = num(k1["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"])
            dmul = num(k1["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"])
            dadd = num(k1["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"])
            peak = num(k1["sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained"])
            summary["executed_fp64"] = {"dfma_per_cycle": dfma, "dmul_per_cycle": dmul, "dadd_per_cycle": dadd,
                                        "flops_per_cycle": 2 * dfma + dmul + dadd, "dfma_peak_per_cycle": peak,
                                        "frac_of_dfma_peak": (2 * dfma + dmul + dadd) / (2 * peak),
                                        "flops_per_launch": (2 * dfma + dmul + dadd) * cyc, "cycles_elapsed_max": cyc}
        except Exception as ex:
            summary["executed_fp64"] = {"error": str(ex)}
        if cfg_id is not None:
            tpath = os.path.join(out, "traffic.json")
            try:
                tj = json.load(open(tpath))
            except Exception:
                tj = {}
            tj.setdefault("configs", {})[cfg_id] = {
                "batch": cfg_batch, "kernel": k0["kernel"][:120], "dram_bytes_per_launch": traffic,
                "executed_fp64_flops_per_launch": (summary.get("executed_fp64") or {}).get("flops_per_launch"),
                "frac_of_dfma_peak_ncu": (summary.get("executed_fp64") or {}).get("frac_of_dfma_peak"),
                "source": f"profiles/{tag}_solve_kernel.json (ncu --set full, hot kernel of bench.py --config {cfg_id})"}
            json.dump(tj, open(tpath, "w"), indent=1)
if os.path.exists(lst):
    rows = [r for r in csv.reader(open(lst)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = defaultdict(list)
    for r in data:
        agg[r[ki].split("(")[0][:70]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    summary["launch_list"] = [{"kernel": k, "launches": len(v), "mean_us": sum(v) / len(v) / 1e3, "share": sum(v) / tot} for k, v in agg.items()]
    with open(os.path.join(out, f"{tag}_launches.csv"), "w") as f:
        f.write(open(lst).read())
json.dump(summary, open(os.path.join(out, f"{tag}_solve_kernel.json"), "w"), indent=1)
print(json.dumps(summary, indent=1)[:3000])
