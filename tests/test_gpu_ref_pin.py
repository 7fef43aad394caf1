"""GPU tier (-m gpu): the CUDA path through the C ABI against numbers computed by the REFERENCE'S OWN SOURCE
(tests/golden/ref_v1.npz = /root/reference/src/QPSolver.cpp compiled unmodified; see tests/test_ref_pin.py and
tests/golden/make_ref_golden.py).  Tolerance 1e-9 relative (north_star) for Ad, Bd, H, f, prediction matrices, bounds."""
import os

import numpy as np
import pytest

from test_ref_pin import case, check_build, rel, TOL

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_v1.npz"))


@pytest.fixture(scope="module")
def lti():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    from mpc_limx_control_b200.lti import LtiContext
    ctx = LtiContext(0)
    yield ctx
    ctx.close()


def test_generic_path_vs_reference_source(ref, lti):
    """mpc_b200_lti_discretize / _build_qp / _update_state (the QPSolver facade's kernels)"""
    n0 = lti.launch_count()
    for nm in ref["case_names"]:
        c, r = case(ref, nm)
        Ad, Bd = lti.discretize(c["Ac"], c["Bc"], c["Ts"])
        assert rel(Ad, r["Ad"]) < TOL and rel(Bd, r["Bd"]) < TOL, nm
        q = lti.build_qp(r["Ad"], r["Bd"], c["Q"], c["R"], c["P"], c["x_min"], c["x_max"], c["u_min"], c["u_max"], c["N"],
                         c["xi0"], c["xi_ref"])
        check_build(q, r, c, nm)
        assert rel(lti.update_state(r["Ad"], r["Bd"], c["xi0"], c["u"]), r["x_next"]) < TOL, nm
    assert lti.launch_count() > n0


@pytest.mark.parametrize("nm", ["tron10a", "tron10b", "tron10stiff", "tron20", "tron50"])
def test_tron1_condense_vs_reference_source(ref, nm):
    """tron1_condense_kernel (closed-form structured condensing, B_aug never formed on the solve path) equals the
    reference's dense buildQPParams on the same Ac, Bc (one model at x0)."""
    import torch
    from mpc_limx_control_b200.engine import Engine
    c, r = case(ref, nm)
    N = c["N"]
    eng = Engine(horizon=N, max_batch=4, device=0, Ts=c["Ts"], ltv=0)
    x0 = torch.from_numpy(c["xi0"][None].copy()).cuda()
    xr = torch.from_numpy(np.ascontiguousarray(c["xi_ref"].T)[None].copy()).cuda()
    feet = torch.from_numpy(c["feet"][None].copy()).cuda()
    d = eng.condense(x0, xr, feet)
    assert rel(d["H"][0], r["H"]) < TOL and rel(d["f"][0], r["f"]) < TOL
    assert not d["B_aug"][0][:13].any()
    assert rel(d["B_aug"][0][13:], r["A_eq"]) < TOL
    assert rel(d["A_aug"][0][13:] @ c["xi0"], r["b_eq"]) < TOL
    eng.close()


def test_demo_closed_loop_vs_reference_source(ref, lti):
    """500 closed-loop steps of the src/qpSolver_test.cpp scenario on the device against the trajectory the reference's
    own buildQPParams/updateState produced (forces bar 1e-4 relative; observed ~1e-9)."""
    c, _ = case(ref, "demo0")
    xs, us = ref["ref_demo_xs"], ref["ref_demo_us"]
    Ad, Bd = lti.discretize(c["Ac"], c["Bc"], 0.01)
    x = np.array([2.0, 0, 0, 0])
    for k in range(500):
        t = (k + np.arange(16)) * 0.01
        xr = np.stack([2 * np.cos(0.5 * t), -np.sin(0.5 * t), 2 * np.sin(0.5 * t), np.cos(0.5 * t)])
        q = lti.build_qp(Ad, Bd, c["Q"], c["R"], c["P"], c["x_min"], c["x_max"], -8.0, 8.0, 15, x, xr)
        U, st, it = lti.qp_solve(q["H"], q["f"], q["A_ineq"], q["lbA_ineq"], q["ubA_ineq"], q["lb"], q["ub"])
        assert st == 0
        x = lti.update_state(Ad, Bd, x, U[:2])
        assert np.abs(U[:2] - us[k]).max() < 1e-4 * max(1.0, np.abs(us[k]).max()), k
        assert np.abs(x - xs[k + 1]).max() < 1e-6, k
