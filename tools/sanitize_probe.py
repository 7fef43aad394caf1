"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
solve (single support, double support overflow class, multi-iteration stress instances), controller-shaped call,
rollout, condense dump, N=20, leg kernels.  usage: compute-sanitizer --tool racecheck python tools/sanitize_probe.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mpc_limx_control_b200 import synth
from mpc_limx_control_b200.engine import Engine, control_host
from mpc_limx_control_b200.leg import LegKinematics

t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for N, B in ((10, 24), (20, 6)):
    d = synth.tron1_batch(3, B, N, 0.02 if N == 10 else 0.005)
    d["x0"][:, [0, 1, 6, 7, 8, 9, 10, 11]] *= 5.0        # stress: several active-face iterations / ADMM
    d["iter"][1] = -1; d["iter"][5] = -1                    # standing instances -> overflow class
    eng = Engine(horizon=N, max_batch=B, Ts=0.02 if N == 10 else 0.005, mu=0.3)
    F, st, it = eng.solve(t(d["x0"]), t(d["x_ref"]), t(d["feet"]), it=t(d["iter"]))
    torch.cuda.synchronize()
    print(f"N={N} status {np.bincount(st.cpu().numpy(), minlength=3)} iters max {int(it.max())}")
    c = eng.condense(t(d["x0"][:2]), t(d["x_ref"][:2]), t(d["feet"][:2]))
    u0, s0, _ = control_host(eng, d["x0"], d["omega_yaw"], d["velocity_x"], d["feet"], it=d["iter"])
    Fh, sh, _ = eng.solve_host(d["x0"], d["x_ref"], d["feet"], it=d["iter"])
    x = t(d["x0"][:8].copy())
    eng.rollout(x, t(d["omega_yaw"][:8]), t(d["velocity_x"][:8]), t(d["iter"][:8]), 6)
    torch.cuda.synchronize()
    eng.close()
lk = LegKinematics()
B = 200
rng = np.random.default_rng(0)
pos = rng.uniform(-1, 1, (B, 3)); quat = rng.normal(size=(B, 4)); q = rng.uniform(-0.5, 0.5, (B, 6))
lk.fk(t(pos), t(quat), t(q), want_jac=True)
qc = t(q.copy())
lk.swing_step(t(pos), t(quat), t(q), t(rng.uniform(-1, 1, (B, 3))), t(rng.integers(0, 5000, B).astype(np.int32)), qc)
lk.grf_to_torque(t(quat), t(q), t(rng.uniform(-1, 1, (B, 6))))
torch.cuda.synchronize()
print("probe done")
