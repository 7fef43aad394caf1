"""Per-phase cycle breakdown of the solve kernel (profiling build, -DMPC_PHASE_TIMING).
Builds a separate library under gpurun_out/ so the product .so is untouched.
usage (on the GPU box): python tools/phase_timing.py [standing]"""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = os.path.join(ROOT, "gpurun_out", "libmpc_b200_timing.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
csrc = os.path.join(ROOT, "mpc_limx_control_b200", "csrc")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DMPC_PHASE_TIMING",
                       *os.environ.get("MPC_EXTRA_FLAGS", "").split(), "-Xcompiler", "-fPIC", "-shared", "-o", out, os.path.join(csrc, "mpc_b200.cu"), os.path.join(csrc, "lti_b200.cu"), os.path.join(csrc, "leg_b200.cu"), os.path.join(csrc, "kf_b200.cu")])
from mpc_limx_control_b200 import _capi, synth
_capi.LIB_PATH = out
import torch
from mpc_limx_control_b200.engine import Engine
standing = "standing" in sys.argv[1:]
N = 50 if "n50" in sys.argv[1:] else (20 if "n20" in sys.argv[1:] else 10)
B = next((int(a) for a in sys.argv[1:] if a.isdigit()), 4096)
d = synth.tron1_batch(1001, B, N, 0.005, standing=standing)
eng = Engine(horizon=N, max_batch=B)
t = {k: torch.from_numpy(d[k]).cuda() for k in ("x0", "x_ref", "feet", "iter")}
L = _capi.lib()
buf = (C.c_ulonglong * 16)()
for _ in range(3):
    eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
torch.cuda.synchronize()
L.mpc_b200_debug_phase_cycles(buf, 1)
F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
torch.cuda.synchronize()
L.mpc_b200_debug_phase_cycles(buf, 0)
names = ["model", "horizon_sums", "free_response", "adjoint(f)", "ufix/grad0", "build_hessian", "rhs", "cholesky", "backward",
         "recover_u", "input_response", "adjoint(g)", "project/check", "tail"]
v = np.array(list(buf), dtype=np.float64)[:14] / B
print(f"B={B} standing={standing} mean iters {it.float().mean().item():.2f}; cycles per instance (thread-0 wall, includes stalls):")
for n_, c in zip(names, v):
    print(f"  {n_:15s} {c:9.0f}  {100 * c / v.sum():5.1f}%")
print(f"  {'total':15s} {v.sum():9.0f}")

# per-CTA timeline of the same launch (globaltimer): how the two rounds of CTAs overlap
def cta_trace(label):
    nb = (B + 3) // 4
    tr = (C.c_ulonglong * (4 * nb))()
    L.mpc_b200_debug_cta_trace(tr, nb)
    tr = np.array(list(tr), dtype=np.float64).reshape(nb, 4)
    t0 = tr[:, 0].min()
    st_, en_, arr = (tr[:, 0] - t0) / 1e3, (tr[:, 1] - t0) / 1e3, (tr[:, 3] - t0) / 1e3
    dur = en_ - st_
    print(f"[{label}] CTAs {nb}: kernel span {en_.max():.1f} us; CTA duration mean {dur.mean():.1f} min {dur.min():.1f} max {dur.max():.1f} us")
    first = st_ < 2.0
    w = arr - st_
    print(f"  first wave: {first.sum()} CTAs, input wait mean {w[first].mean():.1f} max {w[first].max():.1f} us, duration mean {dur[first].mean():.1f} us, end mean {en_[first].mean():.1f} us")
    if (~first).any():
        print(f"  later CTAs: {(~first).sum()}, start mean {st_[~first].mean():.1f} us, input wait mean {w[~first].mean():.1f} max {w[~first].max():.1f}, duration mean {dur[~first].mean():.1f} us, end mean {en_[~first].mean():.1f} max {en_[~first].max():.1f} us")
    q = np.argsort(st_)
    print("  input-arrival time by CTA start order (deciles):", np.round(np.percentile(arr[q][:first.sum()], [0, 10, 25, 50, 75, 90, 100]), 1))
    try:
        tr2 = (C.c_ulonglong * (4 * nb))()
        L.mpc_b200_debug_cta_trace2(tr2, nb)
        tr2 = np.array(list(tr2), dtype=np.float64).reshape(nb, 4)
        for nm, sel in (("first wave", first), ("later CTAs", ~first)):
            if sel.any():
                a = (tr2[sel, 0] - tr[sel, 0]) / 1e3; b_ = (tr2[sel, 1] - tr[sel, 0]) / 1e3; c_ = (tr2[sel, 2] - tr[sel, 0]) / 1e3
                print(f"  {nm}: since CTA start (us, mean/max): copies issued {a.mean():.2f}/{a.max():.2f}, schedule evaluated {b_.mean():.2f}/{b_.max():.2f}, "
                      f"CTA barrier passed {c_.mean():.2f}/{c_.max():.2f}, inputs arrived {w[sel].mean():.2f}/{w[sel].max():.2f}")
    except Exception as ex:
        print("  (no prologue trace:", ex, ")")
    per_sm = np.bincount(tr[:, 2].astype(int), minlength=148)
    print(f"  CTAs per SM: min {per_sm.min()} max {per_sm.max()}")

if N != 10:
    sys.exit(0)
cta_trace("device-resident inputs")
# the same batch through the zero-copy host path (inputs read from pinned host memory by the kernel)
from mpc_limx_control_b200.engine import bind_solve_host
pin = {k: torch.from_numpy(d[k]).pin_memory() for k in ("x0", "x_ref", "feet", "iter")}
Fh = torch.empty((B, N, 6), dtype=torch.float64).pin_memory()
sh = torch.empty(B, dtype=torch.int32).pin_memory(); ih = torch.empty(B, dtype=torch.int32).pin_memory()
hc = bind_solve_host(eng, pin["x0"], pin["x_ref"], pin["feet"], it=pin["iter"], forces=Fh, status=sh, iters=ih)
for _ in range(3):
    hc()
cta_trace("zero-copy host buffers")
