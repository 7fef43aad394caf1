"""CPU tier: the product's generic LTI device source (csrc/lti_core.cuh, the QPSolver path) built for
the host, against the golden demo vectors (src/qpSolver_test.cpp scenario) and the oracle."""
import numpy as np

import emul_lib as E
import oracle_lib as O
import npref as R


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max())


def test_discretize(golden):
    d = R.demo_system()
    Ad, Bd = E.lti_discretize(d["Ac"], d["Bc"], d["Ts"])
    assert rel(Ad, golden["demo_Ad"]) < 1e-12 and rel(Bd, golden["demo_Bd"]) < 1e-12
    rng = np.random.default_rng(0)
    for NX, NU, sc, Ts in [(4, 2, 1.0, 0.01), (13, 6, 3.0, 0.05), (13, 3, 30.0, 0.1), (6, 1, 100.0, 0.2)]:
        Ac = rng.standard_normal((NX, NX)) * sc; Bc = rng.standard_normal((NX, NU))
        Ad, Bd = E.lti_discretize(Ac, Bc, Ts)
        Ad2, Bd2 = O.discretize(Ac, Bc, Ts)
        assert rel(Ad, Ad2) < 1e-9 and rel(Bd, Bd2) < 1e-9


def test_build_matches_golden(golden):
    d = R.demo_system()
    q = E.lti_build(golden["demo_Ad"], golden["demo_Bd"], d["Q"], d["R"], d["P"], d["x_min"], d["x_max"], d["u_min"],
                    d["u_max"], d["N"], np.array([2.0, 0, 0, 0]), R.demo_reference(0, d["Ts"], d["N"]))
    for k, gk in [("A_aug", "demo_A_aug"), ("B_aug", "demo_B_aug"), ("H", "demo_H"), ("f", "demo_f"),
                  ("A_ineq", "demo_A_ineq"), ("lbA_ineq", "demo_lbA"), ("ubA_ineq", "demo_ubA")]:
        assert rel(q[k], golden[gk]) < 1e-9, k
    o = O.build_qp_params(golden["demo_Ad"], golden["demo_Bd"], d["Q"], d["R"], d["P"], d["x_min"], d["x_max"], d["u_min"],
                          d["u_max"], d["N"], np.array([2.0, 0, 0, 0]), R.demo_reference(0, d["Ts"], d["N"]))
    for k in ("A_eq", "b_eq", "lb", "ub"):
        assert rel(q[k], o[k]) < 1e-9, k


def test_build_tron1_literal_model():
    """reference-literal mpcQP model (13 x 3, include/mpcQP.h:139-181) through the generic path"""
    p = O.tron1_defaults()
    Ac, Bc = O.tron1_model_literal(p, np.array([0.1, 0.2, 0.8]), np.array([0.05, 0.1, 0.0]))
    Ad, Bd = E.lti_discretize(Ac, Bc, 0.001)
    Ad2, Bd2 = O.discretize(Ac, Bc, 0.001)
    assert rel(Ad, Ad2) < 1e-12 and rel(Bd, Bd2) < 1e-12
    Q = np.diag(p.q[:]); Rm = 0.1 * np.eye(3); P = 20 * Q
    x0 = np.array([0.01, 0.02, 0.3, 0.1, 0.2, 0.8, 0.0, 0.1, 0.0, 0.3, 0.0, 0.0, -9.8])
    xr = O.tron1_reference(x0, 20, 0.001)
    big = np.full(13, 1e3)
    q = E.lti_build(Ad, Bd, Q, Rm, P, -big, big, -8.0, 8.0, 20, x0, xr)
    o = O.build_qp_params(Ad2, Bd2, Q, Rm, P, -big, big, -8.0, 8.0, 20, x0, xr)
    for k in ("H", "f", "A_aug", "B_aug", "A_ineq", "lbA_ineq", "ubA_ineq"):
        assert rel(q[k], o[k]) < 1e-9, k


def test_dense_qp(golden):
    U, st, it = E.qp_dense(golden["democ_H"], golden["democ_f"], golden["democ_A"], golden["democ_lbA"],
                           golden["democ_ubA"], golden["democ_lb"], golden["democ_ub"])
    assert st == 0
    assert np.abs(U - golden["democ_U"]).max() < 1e-6
    rng = np.random.default_rng(3)
    nsolved = 0
    for _ in range(25):
        n = int(rng.integers(3, 25)); m = int(rng.integers(0, 30))
        M = rng.standard_normal((n, n)); H = M @ M.T + 0.1 * np.eye(n); f = rng.standard_normal(n) * 3
        A = rng.standard_normal((m, n)); xf = rng.standard_normal(n)
        lbA = A @ xf - rng.random(m); ubA = A @ xf + rng.random(m); lb = xf - rng.random(n); ub = xf + rng.random(n)
        ubA[rng.random(m) < 0.3] = O.INFTY; lb[rng.random(n) < 0.3] = -O.INFTY
        e = rng.random(n) < 0.15; lb[e] = ub[e] = xf[e]
        u, info = O.qp_solve(H, f, A, lbA, ubA, lb, ub)
        U, st, it = E.qp_dense(H, f, A if m else None, lbA, ubA, lb, ub)
        if st == 0:
            nsolved += 1
            assert np.abs(U - u).max() / max(1.0, np.abs(u).max()) < 1e-6
        else:
            assert st == 1 and np.isfinite(U).all()          # reported, never silent
            assert np.abs(U - u).max() / max(1.0, np.abs(u).max()) < 5e-2
    assert nsolved >= 20


def test_demo_closed_loop(golden):
    """500 closed-loop steps of the reference demo through the emulated device path"""
    d = R.demo_system()
    Ad, Bd = E.lti_discretize(d["Ac"], d["Bc"], d["Ts"])
    x = np.array([2.0, 0, 0, 0])
    for k in range(500):
        q = E.lti_build(Ad, Bd, d["Q"], d["R"], d["P"], d["x_min"], d["x_max"], d["u_min"], d["u_max"], d["N"], x,
                        R.demo_reference(k, d["Ts"], d["N"]))
        U, st, it = E.qp_dense(q["H"], q["f"], q["A_ineq"], q["lbA_ineq"], q["ubA_ineq"], q["lb"], q["ub"])
        assert st == 0
        x = E.lti_update(Ad, Bd, x, U[:2])
        assert np.abs(U[:2] - golden["demo_us"][k]).max() < 1e-6
        assert np.abs(x - golden["demo_xs"][k + 1]).max() < 1e-6
