"""Where the single-robot host-call latency goes (development tool, run on the GPU box):
kernel time with device-resident inputs, kernel time reading/writing pinned host memory (both event-timed around one
launch), and the wall time of the host call.  python tools/latency_breakdown.py [lib.so]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_limx_control_b200 import _capi, synth
args = sys.argv[1:]
if args and args[0].endswith(".so"):
    _capi.LIB_PATH = os.path.abspath(args[0])
import numpy as np, torch
from mpc_limx_control_b200.engine import Engine, bind_solve_host
N = 10
eng = Engine(horizon=N, max_batch=4096)
L = _capi.lib()
d = synth.tron1_batch(1001, 1, N, 0.005)
stream = torch.cuda.current_stream().cuda_stream
for name, itv in (("walking", int(d["iter"][0])), ("standing", -1)):
    host = {k: torch.from_numpy(d[k][:1].copy()).pin_memory() for k in ("x0", "x_ref", "feet")}
    host["iter"] = torch.full((1,), itv, dtype=torch.int32).pin_memory()
    dev = {k: v.cuda() for k, v in host.items()}
    res = {}
    for where, t in (("device", dev), ("pinned", host)):
        F = torch.empty((1, N, 6), dtype=torch.float64); s = torch.empty(1, dtype=torch.int32); i = torch.empty(1, dtype=torch.int32)
        F, s, i = (F.cuda(), s.cuda(), i.cuda()) if where == "device" else (F.pin_memory(), s.pin_memory(), i.pin_memory())
        p = lambda x: C.c_void_p(x.data_ptr())
        v = []
        for j in range(600):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = L.mpc_b200_tron1_solve_device(eng.h, 1, p(t["x0"]), p(t["x_ref"]), p(t["feet"]), None, p(t["iter"]), p(F), p(s), p(i), C.c_void_p(stream))
            e1.record(); torch.cuda.synchronize()
            assert rc == 0
            if j >= 100:
                v.append(e0.elapsed_time(e1) * 1e3)
        res[where] = np.percentile(v, 50)
    F1 = torch.empty((1, N, 6), dtype=torch.float64).pin_memory(); s1 = torch.empty(1, dtype=torch.int32).pin_memory(); i1 = torch.empty(1, dtype=torch.int32).pin_memory()
    call = bind_solve_host(eng, host["x0"], host["x_ref"], host["feet"], it=host["iter"], forces=F1, status=s1, iters=i1)
    v = []
    for j in range(5200):
        t0 = time.perf_counter(); call()
        if j >= 200:
            v.append((time.perf_counter() - t0) * 1e6)
    print(f"{name}: kernels between events, device-resident inputs {res['device']:.1f} us, pinned host inputs/outputs {res['pinned']:.1f} us "
          f"(device entry = both capacity-class kernels); host call wall p50 {np.percentile(v, 50):.1f} us")
