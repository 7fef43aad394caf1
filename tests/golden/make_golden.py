"""Generates tests/golden/golden_v1.npz from the independent numpy/scipy restatement (npref.py).

The reference ships no golden vectors and cannot be built here, so these
fixtures are the pin for the C oracle: produced by library routines (scipy.linalg.expm,
numpy.linalg.matrix_power, an interior-point QP) that share no code with oracle/mpc_oracle.c.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import npref as R  # noqa: E402
from mpc_limx_control_b200 import synth  # noqa: E402

out = {}

# ---- config 1a: the reference demo (src/qpSolver_test.cpp), first step matrices + closed loop --------
d = R.demo_system()
Ad, Bd = R.discretize(d["Ac"], d["Bc"], d["Ts"])
x = np.array([2.0, 0, 0, 0])
q = R.build_qp_params(Ad, Bd, d["Q"], d["R"], d["P"], d["x_min"], d["x_max"], d["u_min"], d["u_max"], d["N"], x,
                      R.demo_reference(0, d["Ts"], d["N"]))
out.update(demo_Ad=Ad, demo_Bd=Bd, demo_A_aug=q["A_aug"], demo_B_aug=q["B_aug"], demo_H=q["H"], demo_f=q["f"],
           demo_lbA=q["lbA_ineq"], demo_ubA=q["ubA_ineq"], demo_A_ineq=q["A_ineq"])
xs, us = [x.copy()], []
for k in range(500):
    q = R.build_qp_params(Ad, Bd, d["Q"], d["R"], d["P"], d["x_min"], d["x_max"], d["u_min"], d["u_max"], d["N"], x,
                          R.demo_reference(k, d["Ts"], d["N"]))
    U, _ = R.qp_ipm(q["H"], q["f"], q["A_ineq"], q["lbA_ineq"], q["ubA_ineq"], q["lb"], q["ub"])
    u = U[:2]
    x = Ad @ x + Bd @ u
    xs.append(x.copy()); us.append(u.copy())
out.update(demo_xs=np.array(xs), demo_us=np.array(us))

# a constrained variant of the demo (tight input box so the active set is non-trivial)
q = R.build_qp_params(Ad, Bd, d["Q"], d["R"], d["P"], d["x_min"], d["x_max"], -2.0, 2.0, d["N"],
                      np.array([2.0, 0.5, -1.0, 0.2]), R.demo_reference(40, d["Ts"], d["N"]))
U, _ = R.qp_ipm(q["H"], q["f"], q["A_ineq"], q["lbA_ineq"], q["ubA_ineq"], q["lb"], q["ub"])
out.update(democ_H=q["H"], democ_f=q["f"], democ_A=q["A_ineq"], democ_lbA=q["lbA_ineq"], democ_ubA=q["ubA_ineq"],
           democ_lb=q["lb"], democ_ub=q["ub"], democ_U=U)

# ---- TRON1 cases ------------------------------------------------------------------------------------
mu, fmax = 0.5, 2 * R.TRON1_MASS * 9.8
cases = [  # name, N, Ts, ltv, standing, state scale, seed, count
    ("t1b", 10, 0.005, 1, True, 0.0, 1, 1),     # config 1b: standing, zero error
    ("lti", 10, 0.005, 0, False, 1.0, 2, 2),
    ("ltv", 10, 0.005, 1, False, 1.0, 3, 2),
    ("stiff", 10, 0.05, 1, False, 3.0, 4, 2),
    ("short", 10, 0.001, 1, False, 1.0, 5, 2),
    ("n20", 20, 0.02, 1, True, 4.0, 6, 1),
]
names = []
for name, N, Ts, ltv, standing, scale, seed, cnt in cases:
    b = synth.tron1_batch(seed, cnt, N, Ts, standing=standing)
    for i in range(cnt):
        x0 = b["x0"][i].copy()
        x0[[0, 1, 6, 7, 8, 9, 10, 11]] *= scale
        if name == "t1b":   # config 1b of SURVEY.md 8d
            x0 = np.array([0, 0, 0, 0, 0, 0.81181, 0, 0, 0, 0, 0, 0, -9.8])
            feet = np.array([x0[3:6] + R.FOOT_OFFSET_L, x0[3:6] + R.FOOT_OFFSET_R])
            xr = np.repeat(x0[:, None], N + 1, axis=1)
        else:
            feet = b["feet"][i]
            xr = R.tron1_reference(x0, N, Ts, b["omega_yaw"][i], b["velocity_x"][i])
        contact = np.ones((N, 2), np.uint8) if standing else R.contact_schedule(int(b["iter"][i]), N)
        c = R.tron1_condense(x0, xr, feet, N, Ts, ltv)
        A, lbA, ubA, lb, ub = R.tron1_constraints(contact, N, mu, fmax)
        U, _ = R.qp_ipm(c["H"], c["f"], A, lbA, ubA, lb, ub)
        key = f"{name}{i}"
        names.append(key)
        out.update({f"{key}_N": N, f"{key}_Ts": Ts, f"{key}_ltv": ltv, f"{key}_x0": x0, f"{key}_xref": xr,
                    f"{key}_feet": feet, f"{key}_contact": contact, f"{key}_H": c["H"], f"{key}_f": c["f"],
                    f"{key}_A_aug": c["A_aug"], f"{key}_B_aug": c["B_aug"], f"{key}_U": U,
                    f"{key}_Ad0": c["Ad0"], f"{key}_Bd0": c["Bd0"]})
out["tron1_cases"] = np.array(names)

# ---- gait (include/MPCController.h:61-75) -----------------------------------------------------------
rng = np.random.default_rng(7)
iters = np.concatenate([np.arange(0, 2100), rng.integers(0, 2 ** 31 - 1, 4000),
                        np.array([10773998, 10773999, 10774000, 2 ** 31 - 1, 499, 500, 999, 1000])]).astype(np.int64)
g = np.array([R.calculate_gait(int(i)) for i in iters])
out.update(gait_iter=iters, gait_left=g[:, 0].astype(np.int8), gait_right=g[:, 1].astype(np.int8),
           gait_phase=g[:, 2], gait_remain=g[:, 3])

np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
print("wrote golden_v1.npz with", len(out), "arrays")
