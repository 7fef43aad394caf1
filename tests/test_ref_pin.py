"""CPU tier: pins the oracle (and the host build of the product's device source) to numbers computed by the
REFERENCE'S OWN SOURCE -- /root/reference/src/QPSolver.cpp compiled unmodified (oracle/Makefile target `ref`,
fixtures tests/golden/ref_v1.npz, generator tests/golden/make_ref_golden.py).

Rows pinned: a6 discretizeSystem (QPSolver.cpp:21-29), a7 prediction matrices (:36-47, seen through
A_eq = B_aug.bottomRows and b_eq = A_aug.bottomRows x0, :63-64), a8 H and f (:50-60), a9 generic rows (:67-80),
a11 updateState (:108-111).  Tolerance 1e-9 relative (north_star); observed ~1e-13."""
import os

import numpy as np
import pytest

import emul_lib as E
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-9


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max())


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_v1.npz"))


def case(ref, nm):
    cin = {k[len(f"in_{nm}_"):]: ref[k] for k in ref.files if k.startswith(f"in_{nm}_")}
    out = {k[len(f"ref_{nm}_"):]: ref[k] for k in ref.files if k.startswith(f"ref_{nm}_")}
    for k in ("NX", "NU", "N"):
        cin[k] = int(cin[k])
    for k in ("Ts", "u_min", "u_max"):
        cin[k] = float(cin[k])
    return cin, out


def check_build(q, r, c, what):
    NX, N = c["NX"], c["N"]
    for k, rk in [("H", "H"), ("f", "f"), ("A_eq", "A_eq"), ("b_eq", "b_eq"), ("lb", "lb"), ("ub", "ub"), ("A_ineq", "A_ineq"),
                  ("lbA_ineq", "lbA"), ("ubA_ineq", "ubA")]:
        assert rel(q[k], r[rk]) < TOL, (what, k)
    # prediction matrices as such: block row 0 of B_aug is zero, the rest is the reference's A_eq; A_aug x0 rows = b_eq
    assert not q["B_aug"][:NX].any()
    assert rel(q["B_aug"][NX:], r["A_eq"]) < TOL, (what, "B_aug")
    assert rel(q["A_aug"][NX:] @ c["xi0"], r["b_eq"]) < TOL, (what, "A_aug")
    assert np.array_equal(q["A_aug"][:NX], np.eye(NX))
    # unused state rows stay zero with +-INFTY (QPSolver.cpp:71-73)
    for i in range(N):
        assert not r["A_ineq"][2 * i * NX + NX:2 * (i + 1) * NX].any()
        assert (r["lbA"][2 * i * NX + NX:2 * (i + 1) * NX] == -1e20).all() and (r["ubA"][2 * i * NX + NX:2 * (i + 1) * NX] == 1e20).all()


def test_oracle_vs_reference_source(ref):
    """oracle/mpc_oracle.c against the reference's compiled QPSolver on every case"""
    for nm in ref["case_names"]:
        c, r = case(ref, nm)
        Ad, Bd = O.discretize(c["Ac"], c["Bc"], c["Ts"])
        assert rel(Ad, r["Ad"]) < TOL and rel(Bd, r["Bd"]) < TOL, nm
        q = O.build_qp_params(r["Ad"], r["Bd"], c["Q"], c["R"], c["P"], c["x_min"], c["x_max"], c["u_min"], c["u_max"], c["N"],
                              c["xi0"], c["xi_ref"])
        check_build(q, r, c, nm)
        assert rel(O.update_state(r["Ad"], r["Bd"], c["xi0"], c["u"]), r["x_next"]) < TOL, nm


def test_product_source_host_build_vs_reference_source(ref):
    """csrc/lti_core.cuh compiled for the host (tests/emul) against the same reference numbers"""
    for nm in ref["case_names"]:
        c, r = case(ref, nm)
        if c["NX"] + c["NU"] > 64 or c["N"] > 20:
            continue
        Ad, Bd = E.lti_discretize(c["Ac"], c["Bc"], c["Ts"])
        assert rel(Ad, r["Ad"]) < TOL and rel(Bd, r["Bd"]) < TOL, nm
        q = E.lti_build(r["Ad"], r["Bd"], c["Q"], c["R"], c["P"], c["x_min"], c["x_max"], c["u_min"], c["u_max"], c["N"],
                        c["xi0"], c["xi_ref"])
        check_build(q, r, c, nm)
        assert rel(E.lti_update(r["Ad"], r["Bd"], c["xi0"], c["u"]), r["x_next"]) < TOL, nm


@pytest.mark.parametrize("nm", ["tron10a", "tron10b", "tron10stiff", "tron20", "tron50"])
def test_tron1_structured_condensing_vs_reference_source(ref, nm):
    """The TRON1 path never forms B_aug: its closed-form H, f, A_aug, B_aug (one model at x0, the reference's LTI
    structure) must equal what the reference's dense buildQPParams computes from the same Ac, Bc."""
    c, r = case(ref, nm)
    N, Ts = c["N"], c["Ts"]
    x_ref = np.ascontiguousarray(c["xi_ref"].T)
    po = O.tron1_defaults(Ts=Ts, ltv=0)
    o = O.tron1_condense(po, N, c["xi0"], x_ref, c["feet"])
    assert rel(o["H"], r["H"]) < TOL and rel(o["f"], r["f"]) < TOL
    assert rel(o["B_aug"][13:], r["A_eq"]) < TOL and rel(o["A_aug"][13:] @ c["xi0"], r["b_eq"]) < TOL
    if N <= 20:
        e = E.dump(E.default_params(Ts=Ts, ltv=0), N, c["xi0"], x_ref, c["feet"])
        assert rel(e["H"], r["H"]) < TOL and rel(e["f"], r["f"]) < TOL
        assert rel(e["B_aug"][13:], r["A_eq"]) < TOL and rel(e["A_aug"][13:] @ c["xi0"], r["b_eq"]) < TOL
    # the discrete model the kernel uses in closed form equals the reference's expm of the same Ac, Bc
    Ad, Bd = O.discretize(c["Ac"], c["Bc"], Ts)
    assert rel(Ad, r["Ad"]) < TOL and rel(Bd, r["Bd"]) < TOL


def test_demo_closed_loop_vs_reference_source(ref, golden):
    """src/qpSolver_test.cpp:6-50 scenario, 500 steps through the reference's buildQPParams/solveQP/updateState (solver =
    oracle active set behind the qpOASES-shaped shim).  The oracle's own loop and the independent numpy witness
    (golden_v1) must follow it; as written (equality block stacked, row-major read) the first QP is infeasible."""
    assert int(ref["ref_demo_as_written_status"]) == 2
    assert (ref["ref_demo_status"] == 0).all()
    xs, us = ref["ref_demo_xs"], ref["ref_demo_us"]
    assert np.abs(xs - golden["demo_xs"]).max() < 1e-6 and np.abs(us - golden["demo_us"]).max() < 1e-6
    assert rel(ref["ref_demo_H0"], golden["demo_H"]) < TOL and rel(ref["ref_demo_f0"], golden["demo_f"]) < TOL
    # SURVEY.md section 8c pinned scalars
    assert abs(ref["ref_demo_Ad"][0, 1] - 0.009995) < 1e-6 and abs(ref["ref_demo_Bd"][1, 0] - 0.049975) < 1e-6
    assert np.abs(ref["ref_demo_f0"][:2] - [1.092509, -19.409063]).max() < 1e-6
    c, _ = case(ref, "demo0")
    x = np.array([2.0, 0, 0, 0])
    for k in range(60):
        t = (k + np.arange(16)) * 0.01
        xr = np.stack([2 * np.cos(0.5 * t), -np.sin(0.5 * t), 2 * np.sin(0.5 * t), np.cos(0.5 * t)])
        q = O.build_qp_params(ref["ref_demo_Ad"], ref["ref_demo_Bd"], c["Q"], c["R"], c["P"], c["x_min"], c["x_max"], -8.0, 8.0, 15, x, xr)
        u, info = O.qp_solve(q["H"], q["f"], q["A_ineq"], q["lbA_ineq"], q["ubA_ineq"], q["lb"], q["ub"])
        x = O.update_state(ref["ref_demo_Ad"], ref["ref_demo_Bd"], x, u[:2])
        assert np.abs(u[:2] - us[k]).max() < 1e-9 and np.abs(x - xs[k + 1]).max() < 1e-9


@pytest.mark.skipif(not os.path.exists("/root/reference/src/QPSolver.cpp"), reason="reference tree not present on this box")
def test_fixture_is_what_the_reference_source_produces_here():
    """Provenance: rebuild oracle/_ref from the reference where it lies and regenerate -- the committed fixture must be it."""
    import make_ref_golden as G
    fresh = G.generate()
    old = np.load(os.path.join(ROOT, "tests", "golden", "ref_v1.npz"))
    assert set(fresh) == set(old.files)
    for k in old.files:
        if old[k].dtype.kind in "US":
            assert list(old[k]) == list(fresh[k])
        else:
            assert np.array_equal(old[k], np.asarray(fresh[k])), k
