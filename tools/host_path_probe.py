"""Host-buffer call: zero-copy (kernel reads/writes pinned host memory) versus staged copies, and the
device-resident kernel time versus batch size (wave quantisation / latency probe).
usage (GPU box): python tools/host_path_probe.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mpc_limx_control_b200 import synth
from mpc_limx_control_b200.engine import Engine, bind_solve_host, bind_control_host

N, TS = 10, 0.005
eng = Engine(horizon=N, max_batch=65536)


def timeit(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n


print("== host-buffer call, pinned buffers ==")
for B in (1, 8, 64, 512, 4096, 16384, 65536):
    d = synth.tron1_batch(1001, B, N, TS)
    pin = {k: torch.from_numpy(d[k]).pin_memory() for k in ("x0", "x_ref", "feet", "iter", "omega_yaw", "velocity_x")}
    F = torch.empty((B, N, 6), dtype=torch.float64).pin_memory()
    u0 = torch.empty((B, 6), dtype=torch.float64).pin_memory()
    st = torch.empty(B, dtype=torch.int32).pin_memory(); it = torch.empty(B, dtype=torch.int32).pin_memory()
    full = bind_solve_host(eng, pin["x0"], pin["x_ref"], pin["feet"], it=pin["iter"], forces=F, status=st, iters=it)
    ctrl = bind_control_host(eng, pin["x0"], pin["omega_yaw"], pin["velocity_x"], pin["feet"], it=pin["iter"], u0=u0, status=st, iters=it)
    row = [f"B {B:6d}"]
    ref = None
    for mode, name in ((1, "staged"), (2, "zerocopy")):
        eng.set_host_mode(mode)
        n = 2000 if B <= 64 else (300 if B <= 4096 else 40)
        tf = timeit(full, n)
        Fc = F.clone(); sc = st.clone()
        tc = timeit(ctrl, n)
        if ref is None:
            ref = (Fc, u0.clone())
        else:
            assert torch.equal(Fc, ref[0]) and torch.equal(u0, ref[1]) and int((sc != 0).sum()) == 0, "paths disagree"
        row.append(f"{name}: full {tf*1e6:8.1f} us ({B/tf/1e6:6.2f} M/s)  ctrl {tc*1e6:8.1f} us ({B/tc/1e6:6.2f} M/s)")
    print(" | ".join(row))
eng.set_host_mode(0)

print("== device-resident kernel time vs batch ==")
for B in (1, 4, 148, 592, 1184, 2368, 4096, 4736, 9472, 65536):
    d = synth.tron1_batch(1001, B, N, TS)
    t = {k: torch.from_numpy(d[k]).cuda() for k in ("x0", "x_ref", "feet", "iter")}
    F = torch.empty((B, N, 6), dtype=torch.float64, device="cuda")
    st = torch.empty(B, dtype=torch.int32, device="cuda"); it = torch.empty(B, dtype=torch.int32, device="cuda")
    call = eng.bind_solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"], forces=F, status=st, iters=it)
    for _ in range(5):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f"B {B:6d}  {us:8.2f} us/call  {B/us:7.2f} M solves/s")
