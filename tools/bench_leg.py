"""Throughput of the leg kinematics and Kalman kernels (csrc/leg_b200.cu, csrc/kf_b200.cu) against their HBM
roofline.  One JSON line per kernel.  usage (GPU box): python tools/bench_leg.py [B]
(No CPU baseline here: only tests/, smoke() and bench.py may execute the oracle.)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scipy.spatial.transform import Rotation
from mpc_limx_control_b200.leg import LegKinematics

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]; src = "MEASURED_PEAKS.json hbm_gbs"
except Exception:
    peak, src = 6650.0, "fallback 6650 GB/s"
rng = np.random.default_rng(0)
pos = rng.uniform(-1, 1, (B, 3)) + np.array([0, 0, 0.655])
quat = Rotation.from_euler("xyz", rng.uniform([-0.05, -0.05, -np.pi], [0.05, 0.05, np.pi], (B, 3))).as_quat()
q = rng.uniform(-0.05, 0.05, (B, 6)) + np.array([0.0, 0.4, -0.8, 0.0, 0.4, -0.8])
dv = rng.uniform(-0.3, 0.3, (B, 3)); dv[:, 2] = 0
it = rng.integers(0, 10_000_000, B).astype(np.int32)
u0 = rng.uniform(-40, 160, (B, 6))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
d = dict(pos=t(pos), quat=t(quat), q=t(q), dv=t(dv), it=t(it), u0=t(u0), qc=t(q.copy()))
lk = LegKinematics()
tau = torch.empty((B, 6), dtype=torch.float64, device="cuda")


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / n


cases = [
    ("leg_fk_kernel (feet)", 152, lambda: lk.fk(d["pos"], d["quat"], d["q"])),
    ("leg_fk_kernel (feet + Jacobians)", 296, lambda: lk.fk(d["pos"], d["quat"], d["q"], want_jac=True)),
    ("swing_step_kernel", 244, lambda: lk.swing_step(d["pos"], d["quat"], d["q"], d["dv"], d["it"], d["qc"])),
    ("grf_torque_kernel", 176, lambda: lk.grf_to_torque(d["quat"], d["q"], d["u0"], tau)),
]
# Kalman filter: xhat/P resident on the device (read + written every update: 2 x 1248 B), inputs 176 B + 2, odom 104 B
from mpc_limx_control_b200.leg import StateEstimator
Bk = min(B, 1 << 18)
est = StateEstimator(Bk)
kd = dict(quat=d["quat"][:Bk].contiguous(), gyro=t(rng.normal(size=(Bk, 3)) * 0.3), accel=t(rng.normal(size=(Bk, 3)) * 0.5 + np.array([0, 0, 9.81])),
          q=d["q"][:Bk].contiguous(), dq=t(rng.normal(size=(Bk, 6)) * 0.5), contact=t(rng.integers(0, 2, (Bk, 2)).astype(np.uint8)))
odom = torch.empty((Bk, 13), dtype=torch.float64, device="cuda")
s_kf = timeit(lambda: est.update(0.002, kd["quat"], kd["gyro"], kd["accel"], kd["q"], kd["dq"], kd["contact"], odom))
nb = 2 * 1248 + 176 + 2 + 104
print(json.dumps({"kernel": "kf_update_kernel", "B": Bk, "us": s_kf * 1e6, "robots_per_s": Bk / s_kf, "bytes_per_robot": nb,
                  "roofline": {"bound": "hbm", "achieved": nb * Bk / s_kf / 1e9, "peak": peak, "unit": "GB/s", "frac": nb * Bk / s_kf / 1e9 / peak, "peak_source": src}}))
for name, nbytes, fn in cases:
    s = timeit(fn)
    gbs = nbytes * B / s / 1e9
    print(json.dumps({"kernel": name, "B": B, "us": s * 1e6, "robots_per_s": B / s, "bytes_per_robot": nbytes,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": src},
                      "note": "outputs allocated per call by the torch wrapper (cudaMalloc-free caching allocator)"}))
