"""GPU tier: the generic QPSolver path through the C ABI (ctypes) and through the C++ facade
binaries, against the golden demo vectors (reference src/qpSolver_test.cpp scenario) and the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import npref as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "mpc_limx_control_b200", "host", "examples")


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max())


@pytest.fixture(scope="module")
def ctx():
    from mpc_limx_control_b200.lti import LtiContext
    c = LtiContext(0)
    yield c
    c.close()


def test_discretize_build_golden(ctx, golden):
    d = R.demo_system()
    Ad, Bd = ctx.discretize(d["Ac"], d["Bc"], d["Ts"])
    assert rel(Ad, golden["demo_Ad"]) < 1e-9 and rel(Bd, golden["demo_Bd"]) < 1e-9
    q = ctx.build_qp(Ad, Bd, d["Q"], d["R"], d["P"], d["x_min"], d["x_max"], d["u_min"], d["u_max"], d["N"],
                     np.array([2.0, 0, 0, 0]), R.demo_reference(0, d["Ts"], d["N"]))
    for k, gk in [("A_aug", "demo_A_aug"), ("B_aug", "demo_B_aug"), ("H", "demo_H"), ("f", "demo_f"),
                  ("A_ineq", "demo_A_ineq"), ("lbA_ineq", "demo_lbA"), ("ubA_ineq", "demo_ubA")]:
        assert rel(q[k], golden[gk]) < 1e-9, k
    assert np.all(q["lb"] == -8.0) and np.all(q["ub"] == 8.0)


def test_build_batch_and_literal_model(ctx):
    p = O.tron1_defaults()
    Ac, Bc = O.tron1_model_literal(p, np.array([0.1, 0.2, 0.8]), np.array([0.05, 0.1, 0.0]))
    Ad, Bd = ctx.discretize(Ac, Bc, 0.001)
    Ad2, Bd2 = O.discretize(Ac, Bc, 0.001)
    assert rel(Ad, Ad2) < 1e-9 and rel(Bd, Bd2) < 1e-9
    Q = np.diag(p.q[:]); Rm = 0.1 * np.eye(3); P = 20 * Q
    big = np.full(13, 1e3)
    rng = np.random.default_rng(1)
    x0s = np.array([[0.01, 0.02, 0.3, 0.1, 0.2, 0.8, 0.0, 0.1, 0.0, 0.3, 0.0, 0.0, -9.8] + 0 * rng.random(13) for _ in range(3)])
    x0s[:, :12] += rng.uniform(-0.1, 0.1, (3, 12))
    xrs = np.stack([O.tron1_reference(x, 20, 0.001) for x in x0s])
    q = ctx.build_qp(Ad, Bd, Q, Rm, P, -big, big, -8.0, 8.0, 20, x0s, xrs)
    for b in range(3):
        o = O.build_qp_params(Ad2, Bd2, Q, Rm, P, -big, big, -8.0, 8.0, 20, x0s[b], xrs[b])
        for k in ("H", "f", "A_aug", "B_aug", "A_ineq", "lbA_ineq", "ubA_ineq", "A_eq", "b_eq"):
            assert rel(q[k][b], o[k]) < 1e-9, (b, k)


def test_dense_qp_vs_oracle(ctx, golden):
    U, st, it = ctx.qp_solve(golden["democ_H"], golden["democ_f"], golden["democ_A"], golden["democ_lbA"],
                             golden["democ_ubA"], golden["democ_lb"], golden["democ_ub"])
    assert st == 0 and np.abs(U - golden["democ_U"]).max() < 1e-6
    rng = np.random.default_rng(3)
    solved = 0
    for _ in range(15):
        n = int(rng.integers(3, 25)); m = int(rng.integers(0, 30))
        M = rng.standard_normal((n, n)); H = M @ M.T + 0.1 * np.eye(n); f = rng.standard_normal(n) * 3
        A = rng.standard_normal((m, n)); xf = rng.standard_normal(n)
        lbA = A @ xf - rng.random(m); ubA = A @ xf + rng.random(m); lb = xf - rng.random(n); ub = xf + rng.random(n)
        ubA[rng.random(m) < 0.3] = O.INFTY; lb[rng.random(n) < 0.3] = -O.INFTY
        u, info = O.qp_solve(H, f, A, lbA, ubA, lb, ub)
        U, st, it = ctx.qp_solve(H, f, A if m else None, lbA, ubA, lb, ub)
        assert st in (0, 1)
        if st == 0:
            solved += 1
            assert np.abs(U - u).max() / max(1.0, np.abs(u).max()) < 1e-4
            res = np.zeros(4)
    assert solved >= 12


def test_update_state(ctx, golden):
    x = ctx.update_state(golden["demo_Ad"], golden["demo_Bd"], [2.0, 0, 0, 0], golden["demo_us"][0])
    assert np.abs(x - golden["demo_xs"][1]).max() < 1e-12


def test_facade_qp_test_binary(golden):
    """the reference's qp_test program through the C++ QPSolver facade: 500 closed-loop steps"""
    exe = os.path.join(EX, "qp_test")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    r = subprocess.run([exe, "500"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rows = np.array([[float(x) for x in l.split()] for l in r.stdout.strip().splitlines()])
    assert rows.shape == (500, 8)
    assert np.abs(rows[:, 1:3] - golden["demo_us"]).max() < 1e-6
    assert np.abs(rows[:, 3:7] - golden["demo_xs"][1:]).max() < 1e-6
    assert rows[-1, 7] < 0.05


def test_facade_tron1_single_binary(golden):
    """config 1b through MPC::run (controller shim): forces match the golden standing solution"""
    exe = os.path.join(EX, "tron1_single")
    r = subprocess.run([exe, "300"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    vals = lines[0].split()
    f = np.array([float(x) for x in vals[1:7]])
    assert vals[-1] == "1"
    assert np.abs(f - golden["t1b0_U"][:6]).max() < 1e-6
    g = lines[1].split()
    assert g[2] == "1" and g[4] == "0"           # iter 250: left swing, right stance (calculateGait)
    fg = np.array([float(x) for x in g[6:12]])
    assert np.all(fg[:3] == 0.0) and fg[5] > 0.0  # swing foot carries no force
    # the rest of MPC::run: FK, swing-leg targets and stance torques (csrc/leg_b200.cu) against the oracle
    import oracle_lib as O
    mo, po = O.leg_defaults()
    q = np.array([0.0, 0.4, -0.8, 0.0, 0.4, -0.8], np.float32).astype(np.float64)
    pos, quat = np.array([0.0, 0.0, 0.81181]), np.array([0.0, 0.0, 0.0, 1.0])
    ref = O.swing_step(mo, po, pos, quat, q, [0.0, 0.0, 0.0], 250, np.zeros(6))
    tau = np.array([float(x) for x in lines[2].split()[2:8]])
    fk = np.array([float(x) for x in lines[3].split()[2:8]])
    sw = lines[4].split()
    nxt = np.array([float(x) for x in sw[2:5]]); cq = np.array([float(x) for x in sw[6:12]])
    assert ref["leg"] == 0 and np.abs(fk - ref["feet"].reshape(-1)).max() < 1e-12 and np.abs(nxt - ref["next_foot"]).max() < 1e-12
    assert np.abs(cq[:3] - ref["q_cmd"][:3].astype(np.float32)).max() < 1e-6 and np.all(cq[3:] == 0.0)    # cmd.q is float
    assert int(sw[-1]) == ref["ik_iters"] and abs(float(sw[-3]) - ref["ik_err"]) < 1e-6 * max(1.0, ref["ik_err"])
    assert np.abs(tau - O.grf_to_torque(mo, quat, q, fg)).max() < 1e-9 and np.all(tau[:3] == 0.0) and np.abs(tau[3:]).max() > 0.1
    # the Kalman estimator shim (host/stateEstimator.h) standing still: height = leg length + foot radius, zero velocity,
    # and the controller driven by it produces a certified, upward, near-symmetric standing force
    by = lambda pre: next(l for l in lines if l.startswith(pre))
    # the reference's own call shape: PinocchioKinematics shim + mpcQP(state, pos, vel, rpy, omega, quat, kin, leg)
    # (include/mpcQP.h:10,35-119; include/pinocchio_kinematics.h:30-43,61-149,153-157)
    mc = by("mpcQP_ctor").split()
    pos2 = np.array([0.0, 0.0, 0.655])
    feet_c = np.array([float(x) for x in mc[2:8]]); u_c = np.array([float(x) for x in mc[9:15]])
    fk_ref = np.concatenate([O.leg_fk(mo, l, pos2, quat, q[3 * l:3 * l + 3], want_jac=False) for l in (0, 1)])
    assert np.abs(feet_c - fk_ref).max() < 1e-11 and mc[-1] == "1"
    x0 = np.array([0, 0, 0, 0, 0, 0.655, 0, 0, 0, 0.2, 0, 0, -9.8])
    pt = O.tron1_defaults(Ts=0.005)
    xr = O.tron1_reference(x0, 10, 0.005, 0.1, 0.5).T.copy()
    contact = np.tile(np.array([1, 0], np.uint8), (10, 1))           # leg = 0: the left foot is the support foot
    Fo, so, _ = O.tron1_solve_batch(pt, 10, x0[None], xr[None], fk_ref.reshape(1, 2, 3), contact[None], nthreads=1)
    assert so[0] == 0 and np.abs(u_c - Fo[0, 0]).max() < 1e-4 * max(1.0, np.abs(Fo[0, 0]).max()) and np.all(u_c[3:] == 0.0)
    target = fk_ref[:3] + np.array([0.02, -0.01, 0.03])
    for mode in (0, 1):
        ik = by(f"ik_mode{mode}").split()
        po.ik_mode = mode
        qo, eo, io = O.leg_ik(mo, po, 0, pos2, quat, target, q[:3])
        assert np.abs(np.array([float(x) for x in ik[2:5]]) - qo).max() < 1e-8 and int(ik[-1]) == io and abs(float(ik[6]) - eo) < 1e-8
    po.ik_mode = 0
    assert by("missing_frame").split()[1:] == ["0.0", "0.0", "0.0"]
    est = by("estimator pos").split()
    feet0 = O.leg_fk(mo, 0, [0, 0, 0], quat, q[:3], want_jac=False)
    assert abs(float(est[4]) - (-feet0[2] + 0.02)) < 2e-3 and max(abs(float(v)) for v in est[6:9]) < 1e-3
    assert abs(float(est[12])) < 2e-3          # left foot on the ground plane
    ef = by("estimator-driven forces").split()
    assert ef[-1] == "1" and float(ef[4]) > 1.0 and float(ef[7]) > 1.0
    assert by("latency_us")
