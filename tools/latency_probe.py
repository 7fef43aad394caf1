"""Single-robot latency of a library variant through the host entry (B = 1, pinned buffers):
python tools/latency_probe.py [lib.so] [calls]   -- prints p50/p99 for the walking and the standing robot (development tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_limx_control_b200 import _capi, synth
args = sys.argv[1:]
if args and args[0].endswith(".so"):
    _capi.LIB_PATH = os.path.abspath(args[0]); args = args[1:]
import numpy as np, torch
from mpc_limx_control_b200.engine import Engine, bind_solve_host
calls = int(args[0]) if args else 5000
N = 10
eng = Engine(horizon=N, max_batch=4096)
d = synth.tron1_batch(1001, 1, N, 0.005)
one = {k: torch.from_numpy(d[k][:1].copy()).pin_memory() for k in ("x0", "x_ref", "feet", "iter")}
F1 = torch.empty((1, N, 6), dtype=torch.float64).pin_memory()
s1 = torch.empty(1, dtype=torch.int32).pin_memory(); i1 = torch.empty(1, dtype=torch.int32).pin_memory()
out = []
for name, it in (("walking", one["iter"]), ("standing", torch.full((1,), -1, dtype=torch.int32).pin_memory())):
    call = bind_solve_host(eng, one["x0"], one["x_ref"], one["feet"], it=it, forces=F1, status=s1, iters=i1)
    v = []
    for j in range(calls + 200):
        t0 = time.perf_counter(); call()
        if j >= 200:
            v.append(time.perf_counter() - t0)
    v = np.array(v) * 1e6
    out.append(f"{name} p50 {np.percentile(v, 50):.1f} p99 {np.percentile(v, 99):.1f} us (status {int(s1[0])})")
print(os.path.basename(_capi.LIB_PATH), " | ".join(out))
