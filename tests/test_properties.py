"""CPU tier: property tests (hypothesis) of the emulated device source on randomly drawn problems.
Every certified solution must satisfy the solver-independent KKT measure (oracle's natural residual with
the exact pyramid projection) on the oracle's dense H, f -- whatever the parameters."""
import numpy as np
from hypothesis import given, settings, strategies as st, HealthCheck

import emul_lib as E
import oracle_lib as O
from mpc_limx_control_b200 import synth


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 10 ** 6), Ts=st.sampled_from([0.001, 0.005, 0.02, 0.05]), mu=st.floats(0.1, 1.0),
       fscale=st.floats(0.2, 3.0), scale=st.floats(0.0, 6.0), ltv=st.integers(0, 1), p_on=st.floats(0.2, 1.0),
       qscale=st.floats(0.1, 10.0), r=st.floats(0.01, 1.0))
def test_certified_solutions_satisfy_kkt(seed, Ts, mu, fscale, scale, ltv, p_on, qscale, r):
    N = 10
    d = synth.tron1_batch(seed, 1, N, Ts)
    x0 = d["x0"][0].copy(); x0[[0, 1, 6, 7, 8, 9, 10, 11]] *= scale
    rng = np.random.default_rng(seed)
    contact = (rng.random((N, 2)) < p_on).astype(np.uint8)
    fmax = fscale * 9.585 * 9.8
    q = np.array([1, 1, 10, 100, 100, 100, 50, 50, 50, 100, 100, 100, 0.1]) * qscale
    pe = E.default_params(Ts=Ts, mu=mu, f_max=fmax, ltv=ltv, q=q, r=r)
    po = O.tron1_defaults(Ts=Ts, mu=mu, f_max=fmax, ltv=ltv, q=q, r=r)
    F, status, it = E.solve(pe, N, x0, d["x_ref"][0], d["feet"][0], contact)
    assert status in (0, 1) and np.isfinite(F).all()
    F3 = F.reshape(N, 2, 3)
    # feasibility always (also for iteration-capped returns)
    assert np.all(F3[contact == 0] == 0.0)
    assert (np.abs(F3[..., 0]) <= mu * F3[..., 2] + 1e-8).all() and (np.abs(F3[..., 1]) <= mu * F3[..., 2] + 1e-8).all()
    assert (F3[..., 2] >= -1e-10).all() and (F3[..., 2] <= fmax + 1e-8).all()
    if status == 0:
        c = O.tron1_condense(po, N, x0, d["x_ref"][0], d["feet"][0], want_pred=False)
        res = O.tron1_natural_residual(po, N, c["H"], c["f"], contact, F)
        assert res <= 1e-6 * max(1.0, np.abs(F).max())


@settings(max_examples=25, deadline=None)
@given(it=st.integers(0, 2 ** 31 - 1 - 60))
def test_gait_schedule_bit_exact(it):
    pe = E.default_params()
    assert np.array_equal(E.gait_contact(pe, it, 10), O.contact_schedule(it, 10))
