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


# ---- leg kinematics and Kalman filter: invariances of the product's kernel mathematics ----------
import ctypes as _C
from scipy.spatial.transform import Rotation as _Rot
from mpc_limx_control_b200 import _capi as _capi


def _leg_defaults():
    m, p, k = _capi.LegModel(), _capi.SwingParams(), _capi.KfParams()
    L = _capi.lib()
    L.mpc_b200_leg_default_model(_C.byref(m)); L.mpc_b200_swing_default_params(_C.byref(p)); L.mpc_b200_kf_default_params(_C.byref(k))
    return m, p, k


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 10 ** 6))
def test_leg_kinematics_invariances(seed):
    m, p, _ = _leg_defaults()
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-2, 2, 3); q = rng.uniform(-0.8, 0.8, 6)
    r1 = _Rot.from_euler("xyz", rng.uniform(-1, 1, 3)); r2 = _Rot.from_euler("xyz", rng.uniform(-3, 3, 3))
    feet, jac = E.leg_fk(m, pos, r1.as_quat(), q)
    # rigid motion of the base moves the feet rigidly and rotates the Jacobians
    shift = rng.uniform(-1, 1, 3)
    feet2, jac2 = E.leg_fk(m, r2.apply(pos) + shift, (r2 * r1).as_quat(), q)
    assert np.abs(feet2 - (r2.apply(feet) + shift)).max() < 1e-12
    assert np.abs(jac2 - np.einsum("ij,ljk->lik", r2.as_matrix(), jac)).max() < 1e-12
    # the quaternion is normalised: scaling it changes nothing; -q is the same rotation
    feet3, _ = E.leg_fk(m, pos, -3.7 * r1.as_quat(), q)
    assert np.abs(feet3 - feet).max() < 1e-12
    # torque is linear in the force and the two legs are decoupled
    f = rng.uniform(-100, 100, 6)
    tau = E.grf_to_torque(m, r1.as_quat(), q, f)
    fl = f.copy(); fl[3:] = 0
    assert np.abs(E.grf_to_torque(m, r1.as_quat(), q, 2.5 * f) - 2.5 * tau).max() < 1e-10
    assert np.abs(E.grf_to_torque(m, r1.as_quat(), q, fl)[:3] - tau[:3]).max() < 1e-12 and np.all(E.grf_to_torque(m, r1.as_quat(), q, fl)[3:] == 0)
    # an IK target at the current foot position is a fixed point: zero iterations, joints unchanged
    p.ik_max_iter = 10
    it = int(rng.integers(0, 10 ** 6))
    out = E.swing_step(m, p, pos, r1.as_quat(), q, rng.uniform(-1, 1, 3), it, q)
    assert out["ik_iters"] <= 10 and np.isfinite(out["q_cmd"]).all()
    st_leg = 1 - out["leg"]
    assert np.array_equal(out["q_cmd"][3 * st_leg:3 * st_leg + 3], q[3 * st_leg:3 * st_leg + 3])


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 10 ** 6), dt=st.sampled_from([0.0005, 0.001, 0.002, 0.005, 0.01]))
def test_kalman_update_properties(seed, dt):
    m, _, k = _leg_defaults()
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(12, 12)); P = A @ A.T * 10.0 ** rng.uniform(-4, 2) + 1e-6 * np.eye(12)
    x = rng.normal(size=12)
    quat = _Rot.from_euler("xyz", rng.uniform(-0.5, 0.5, 3)).as_quat()
    u = dict(gyro=rng.normal(size=3), accel=rng.normal(size=3) + [0, 0, 9.81], q=rng.uniform(-0.5, 0.5, 6), dq=rng.normal(size=6))
    contact = rng.integers(0, 2, 2)
    x1, P1, od = E.kf_update(k, m, dt, quat, u["gyro"], u["accel"], u["q"], u["dq"], contact, x, P)
    assert np.isfinite(x1).all() and np.array_equal(P1, P1.T)
    assert np.linalg.eigvalsh(P1).min() > -1e-9 * max(1.0, np.abs(P1).max())          # stays positive semi-definite
    # a measurement update cannot increase uncertainty beyond the prediction (before the reference's 2x2 reset rescales)
    a = np.eye(12); a[0:3, 3:6] = dt * np.eye(3)
    pm_diag = np.diag(a @ P @ a.T) + np.r_[np.full(3, dt / 20 * 0.02), np.full(3, dt * 9.81 / 20 * 0.02), np.full(6, dt * 0.002 * 100)]
    assert (np.diag(P1)[2:] <= pm_diag[2:] * (1 + 1e-9) + 1e-15).all()
    # translating the state estimate in x/y (base and feet together) translates the result: the filter has no absolute xy
    sh = np.zeros(12); sh[[0, 6, 9]] = 0.37; sh[[1, 7, 10]] = -1.2
    x2, P2, _ = E.kf_update(k, m, dt, quat, u["gyro"], u["accel"], u["q"], u["dq"], contact, x + sh, P)
    assert np.abs(x2 - (x1 + sh)).max() < 1e-9 * max(1.0, np.abs(x1).max()) and np.abs(P2 - P1).max() < 1e-9 * max(1.0, np.abs(P1).max())
