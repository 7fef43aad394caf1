"""CPU tier: pins the C oracle against the golden fixtures produced by the independent
numpy/scipy restatement (tests/golden/make_golden.py).  Tolerances: 1e-9 relative (FP64) for
matrices as BASELINE.json's north_star states, 1e-4 relative for QP solutions (we assert tighter)."""
import numpy as np
import pytest

import oracle_lib as O
import npref as R


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max())


def test_expm_matches_scipy():
    import scipy.linalg as sla
    rng = np.random.default_rng(0)
    for n, sc in [(6, 0.01), (6, 0.2), (6, 0.8), (6, 1.5), (19, 3.0), (19, 30.0)]:
        A = rng.standard_normal((n, n)) * sc
        assert rel(O.expm(A), sla.expm(A)) < 1e-12


def test_demo_discretize_and_condense(golden):
    d = R.demo_system()
    Ad, Bd = O.discretize(d["Ac"], d["Bc"], d["Ts"])
    assert rel(Ad, golden["demo_Ad"]) < 1e-12 and rel(Bd, golden["demo_Bd"]) < 1e-12
    # SURVEY.md 8c pinned scalars
    assert abs(Ad[0, 1] - 0.009995) < 1e-6 and abs(Bd[1, 0] - 0.049975) < 1e-6
    q = O.build_qp_params(Ad, Bd, d["Q"], d["R"], d["P"], d["x_min"], d["x_max"], d["u_min"], d["u_max"], d["N"],
                          np.array([2.0, 0, 0, 0]), R.demo_reference(0, d["Ts"], d["N"]))
    for k, gk in [("A_aug", "demo_A_aug"), ("B_aug", "demo_B_aug"), ("H", "demo_H"), ("f", "demo_f"),
                  ("A_ineq", "demo_A_ineq"), ("lbA_ineq", "demo_lbA"), ("ubA_ineq", "demo_ubA")]:
        assert rel(q[k], golden[gk]) < 1e-9, k
    assert np.allclose(q["f"][:2], [1.092509, -19.409063], atol=1e-6)
    ev = np.linalg.eigvalsh(q["H"])
    assert abs(ev[0] - 0.2063) < 1e-3 and abs(ev[-1] - 9.8599) < 1e-3


def test_demo_closed_loop(golden):
    """500 closed-loop steps of src/qpSolver_test.cpp with the oracle's active-set QP."""
    d = R.demo_system()
    Ad, Bd = O.discretize(d["Ac"], d["Bc"], d["Ts"])
    x = np.array([2.0, 0, 0, 0])
    for k in range(500):
        q = O.build_qp_params(Ad, Bd, d["Q"], d["R"], d["P"], d["x_min"], d["x_max"], d["u_min"], d["u_max"], d["N"],
                              x, R.demo_reference(k, d["Ts"], d["N"]))
        U, info = O.qp_solve(q["H"], q["f"], q["A_ineq"], q["lbA_ineq"], q["ubA_ineq"], q["lb"], q["ub"])
        assert info["status"] == 0 and info["kkt"].max() < 1e-6
        x = O.update_state(Ad, Bd, x, U[:2])
        assert np.abs(U[:2] - golden["demo_us"][k]).max() < 1e-6
        assert np.abs(x - golden["demo_xs"][k + 1]).max() < 1e-6
    # tracking error settles
    ref = R.demo_reference(499, d["Ts"], d["N"])[:, 1]
    assert np.hypot(x[0] - ref[0], x[2] - ref[2]) < 0.05


def test_constrained_demo_qp(golden):
    U, info = O.qp_solve(golden["democ_H"], golden["democ_f"], golden["democ_A"], golden["democ_lbA"],
                         golden["democ_ubA"], golden["democ_lb"], golden["democ_ub"])
    assert info["status"] == 0 and info["kkt"].max() < 1e-8
    assert np.abs(U - golden["democ_U"]).max() < 1e-6
    assert (np.abs(U) > 2.0 - 1e-9).sum() > 0   # the box is really active


def test_random_qps_vs_ipm():
    rng = np.random.default_rng(3)
    for _ in range(25):
        n = int(rng.integers(3, 25)); m = int(rng.integers(0, 30))
        M = rng.standard_normal((n, n)); H = M @ M.T + 0.1 * np.eye(n); f = rng.standard_normal(n) * 3
        A = rng.standard_normal((m, n)); xf = rng.standard_normal(n)
        lbA = A @ xf - rng.random(m); ubA = A @ xf + rng.random(m); lb = xf - rng.random(n); ub = xf + rng.random(n)
        ubA[rng.random(m) < 0.3] = O.INFTY; lb[rng.random(n) < 0.3] = -O.INFTY
        e = rng.random(n) < 0.15; lb[e] = ub[e] = xf[e]
        u, info = O.qp_solve(H, f, A, lbA, ubA, lb, ub)
        u2, _ = R.qp_ipm(H, f, A, lbA, ubA, lb, ub)
        assert info["status"] == 0 and info["kkt"].max() < 1e-8
        assert np.abs(u - u2).max() < 1e-6


def test_infeasible_qp_reports_status():
    H = np.eye(2); f = np.zeros(2)
    A = np.array([[1.0, 0.0]]); u, info = O.qp_solve(H, f, A, np.array([2.0]), np.array([O.INFTY]), np.array([-1.0, -1.0]), np.array([1.0, 1.0]))
    assert info["status"] == 2


def test_gait_bit_exact(golden):
    for it, l, r, ph, rem in zip(golden["gait_iter"], golden["gait_left"], golden["gait_right"], golden["gait_phase"], golden["gait_remain"]):
        a = O.calculate_gait(int(it))
        assert a[0] == l and a[1] == r and a[2] == ph and a[3] == rem, it
    # float semantics matter: a pure-double evaluation disagrees at this iteration
    it = 10773999
    dbl_phase = np.fmod(it * 0.001, 1.0)
    assert (dbl_phase < 0.5) != (O.calculate_gait(it)[0] == 1)


def test_tron1_cases(golden):
    for key in golden["tron1_cases"]:
        N = int(golden[f"{key}_N"]); Ts = float(golden[f"{key}_Ts"]); ltv = int(golden[f"{key}_ltv"])
        p = O.tron1_defaults(Ts=Ts, ltv=ltv)
        x0 = golden[f"{key}_x0"]; xr = golden[f"{key}_xref"].T.copy(); feet = golden[f"{key}_feet"]
        c = O.tron1_condense(p, N, x0, xr, feet)
        for k in ("H", "f", "A_aug", "B_aug"):
            assert rel(c[k], golden[f"{key}_{k}"]) < 1e-9, (key, k)
        contact = golden[f"{key}_contact"]
        A, lbA, ubA, lb, ub = O.tron1_constraints(p, N, contact)
        U, info = O.qp_solve(c["H"], c["f"], A, lbA, ubA, lb, ub)
        assert info["status"] == 0 and info["kkt"].max() < 1e-6, key
        assert np.abs(U - golden[f"{key}_U"]).max() / max(1.0, np.abs(U).max()) < 1e-5, key
        assert O.tron1_natural_residual(p, N, c["H"], c["f"], contact, U) < 1e-6
        # swing feet carry exactly zero force
        assert np.all(U.reshape(N, 2, 3)[contact == 0] == 0.0)


def test_tron1_model_properties():
    p = O.tron1_defaults()
    pos = np.array([0.1, -0.2, 0.8]); feet = np.array([[0.05, -0.3, 0.0], [0.12, -0.1, 0.0]])
    Ac, Bc = O.tron1_model(p, 0.7, pos, feet)
    Ac2, Bc2 = R.tron1_model(0.7, pos, feet)
    assert rel(Ac, Ac2) < 1e-14 and rel(Bc, Bc2) < 1e-13
    assert np.abs(Ac @ Ac @ Ac).max() == 0 and np.abs(Ac @ Ac @ Bc).max() == 0   # nilpotent SRBD
    Ad, Bd = O.discretize(Ac, Bc, 0.005)
    assert rel(Ad, np.eye(13) + 0.005 * Ac + 0.005 ** 2 / 2 * Ac @ Ac) < 1e-15
    assert rel(Bd, 0.005 * Bc + 0.005 ** 2 / 2 * Ac @ Bc) < 1e-14
    # reference-literal switch (include/mpcQP.h:139-181): the divergence is explicit and testable
    Al, Bl = O.tron1_model_literal(p, pos, feet[0])
    Al2, Bl2 = R.tron1_model_literal(pos, feet[0])
    assert np.array_equal(Al, Al2) and np.array_equal(Bl, Bl2)
    assert Al[11, 12] == -1.0 and Ac[11, 12] == 1.0 and Bl[9, 0] == -p.mass and abs(Bc[9, 0] - 1 / p.mass) < 1e-16


def test_tron1_reference_generator():
    x0 = np.arange(13) * 0.1
    xr = O.tron1_reference(x0, 10, 0.005, 0.1, 0.5)
    assert rel(xr, R.tron1_reference(x0, 10, 0.005, 0.1, 0.5)) < 1e-15
    assert xr[9, 0] == x0[9] and xr[9, 1] == 0.5 and xr[12, 3] == -9.8


def test_standing_symmetry(golden):
    """config 1b: zero tracking error, feet placed symmetrically about the base -> both feet carry
    (almost) the same vertical force; exact mirror symmetry is broken only by the products of
    inertia Ixy, Iyz of include/mpcQP.h:20-22."""
    U = golden["t1b0_U"].reshape(10, 2, 3)
    assert np.all(U[:, :, 2] > 0)
    assert np.allclose(U[:, 0, 2], U[:, 1, 2], rtol=1e-5)
    assert np.abs(U[:, :, 1]).max() < 1e-4
