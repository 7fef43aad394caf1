"""Device-resident throughput of a library variant: python tools/variant_probe.py <lib.so> [n50|n20] [standing] [B ...]
(development tool: compares builds of csrc/ with different compile-time choices on the same inputs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_limx_control_b200 import _capi, synth
if len(sys.argv) > 1 and sys.argv[1].endswith(".so"):
    _capi.LIB_PATH = os.path.abspath(sys.argv[1]); args = sys.argv[2:]
else:
    args = sys.argv[1:]
import torch
from mpc_limx_control_b200.engine import Engine
N, TS = (50 if "n50" in args else 20 if "n20" in args else 10), 0.005
standing = "standing" in args
pipelined = "pipelined" in args
Bs = [int(a) for a in args if a.isdigit()] or [4096, 65536]
eng = Engine(horizon=N, max_batch=max(Bs))
out = []
for B in Bs:
    d = synth.tron1_batch(1001, B, N, TS, standing=standing)
    t = {k: torch.from_numpy(d[k]).cuda() for k in ("x0", "x_ref", "feet", "iter")}
    F = torch.empty((B, N, 6), dtype=torch.float64, device="cuda")
    st = torch.empty(B, dtype=torch.int32, device="cuda"); it = torch.empty(B, dtype=torch.int32, device="cuda")
    call = eng.bind_solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"], forces=F, status=st, iters=it, pipelined=pipelined)
    for _ in range(10):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = (300 if B <= 8192 else 50) if N == 10 else 5
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); e0.record()
        for _ in range(n):
            call()
        if pipelined:
            eng.join()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    out.append(f"B {B}: {best:.2f} us {B/best:.1f} M/s (bad {int((st != 0).sum())})")
print(os.path.basename(_capi.LIB_PATH), f"N={N} standing={standing} pipelined={pipelined}", " | ".join(out))
