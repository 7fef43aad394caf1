"""Opcode histogram and resource usage per kernel of the built library -> profiles/sass_summary.txt
(evidence for the SASS claims in DESIGN.md: UBLKCP / SYNCS = 1-D TMA bulk copies + mbarriers, DMMA = FP64 tensor-core MMA,
LDGSTS = cp.async, MUFU.RCP64H / RSQ64H, register counts, spills).  Runs here (no GPU needed): cuobjdump only.
usage: python tools/sass_summary.py [lib.so] [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mpc_limx_control_b200", "csrc", "libmpc_b200.so")
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "sass_summary.txt")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n

usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in line:
        usage[cur] = line.strip()
        cur = None

kern = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1); kern[name] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Za-z0-9_]+)*)", line)
    if m and name:
        kern[name][m.group(1)] += 1

INTEREST = ["DFMA", "DMUL", "DADD", "DMMA", "MUFU", "UBLKCP", "SYNCS", "LDGSTS", "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "BAR", "ATOMG", "RED", "UTMALDG", "HMMA"]
with open(out, "w") as f:
    f.write(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts (static) and resource usage per kernel; tools/sass_summary.py\n")
    total = collections.Counter()
    for k, c in kern.items():
        base = collections.Counter()
        for op, n in c.items():
            base[op.split(".")[0]] += n
            total[op.split(".")[0]] += n
        f.write(f"\n== {demangle(k)[:200]}\n   {usage.get(k, '')}\n   instructions {sum(c.values())}: ")
        f.write(" ".join(f"{op}={base[op]}" for op in INTEREST if base[op]) + "\n")
        special = {op: n for op, n in c.items() if op.startswith(("DMMA", "MUFU.RCP64H", "MUFU.RSQ64H", "UBLKCP", "SYNCS", "LDGSTS", "BAR.RED", "BAR.ARV"))}
        if special:
            f.write("   " + " ".join(f"{op}={n}" for op, n in sorted(special.items())) + "\n")
    f.write("\n== whole library: " + " ".join(f"{op}={total[op]}" for op in INTEREST if total[op]) + "\n")
print(open(out).read()[:3000])
