"""CPU tier for the leg kinematics: the oracle (oracle/leg_oracle.c) is pinned against an
independent numpy/scipy witness and finite differences, and the product's kernel mathematics (csrc/leg_core.cuh,
compiled for the host in tests/emul) is checked against the oracle.  Parity with the reference itself is unpinned
for this block (external URDF, no Pinocchio): see oracle/leg_oracle.h."""
import ctypes as C

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import emul_lib as E
import oracle_lib as O
from mpc_limx_control_b200 import _capi


def product_defaults():
    m, p = _capi.LegModel(), _capi.SwingParams()
    L = _capi.lib()
    assert L.mpc_b200_leg_default_model(C.byref(m)) == 0 and L.mpc_b200_swing_default_params(C.byref(p)) == 0
    return m, p


def np_fk(model, leg, pos, quat, q):
    """numpy/scipy witness: chain of scipy rotations (different code path from both the oracle and the product)."""
    off = np.array(model.offset).reshape(2, 5, 3)[leg]
    ax = np.array(model.axis).reshape(2, 3, 3)[leg]
    Rb = Rotation.from_quat(np.asarray(quat) / np.linalg.norm(quat))   # scipy: [x, y, z, w]
    R = Rb
    p = np.asarray(pos, float) + Rb.apply(off[0])
    for k in range(3):
        R = R * Rotation.from_rotvec(ax[k] * q[k])
        nxt = off[k + 1] if k < 2 else off[3] + off[4]
        p = p + R.apply(nxt)
    return p


def rand_state(rng):
    pos = rng.uniform(-1, 1, 3) + np.array([0, 0, 0.8])
    quat = Rotation.from_euler("xyz", rng.uniform([-0.3, -0.3, -np.pi], [0.3, 0.3, np.pi])).as_quat()
    q = rng.uniform(-0.6, 0.6, 6)
    return pos, quat, q


def test_defaults_match_reference_literals():
    mo, po = O.leg_defaults()
    mp, pp = product_defaults()
    assert list(mo.offset) == list(mp.offset) and list(mo.axis) == list(mp.axis)
    for f, _ in O.SwingParams._fields_:
        a, b = getattr(po, f), getattr(pp, f)
        assert (list(a) == list(b)) if hasattr(a, "__len__") else a == b, f
    # include/MPCParam.h:64-73 and SURVEY.md 8c: nominal foot offsets
    assert np.allclose(list(pp.foot_offset_left), [-0.02644, -0.105, -0.81181], atol=1e-12)
    assert np.allclose(list(pp.foot_offset_right), [-0.02644, 0.105, -0.81181], atol=1e-12)
    assert pp.ik_max_iter == 10 and pp.ik_tol == 1e-3 and pp.ik_dt == 0.1 and pp.ik_damp == 1e-6


def test_zero_pose_reproduces_static_foot_offsets():
    """q = 0, identity base: the contact points sit at static_foot_offset_{left,right} (include/MPCParam.h:64-73)."""
    m, p = O.leg_defaults()
    for leg, off in ((0, p.foot_offset_left), (1, p.foot_offset_right)):
        pt = O.leg_fk(m, leg, [0, 0, 0], [0, 0, 0, 1], [0, 0, 0], want_jac=False)
        assert np.abs(pt - np.array(list(off))).max() < 1e-15


def test_oracle_fk_matches_numpy_witness_and_fd_jacobian():
    m, _ = O.leg_defaults()
    rng = np.random.default_rng(0)
    for _ in range(50):
        pos, quat, q = rand_state(rng)
        for leg in (0, 1):
            p, J = O.leg_fk(m, leg, pos, quat, q[3 * leg:3 * leg + 3])
            assert np.abs(p - np_fk(m, leg, pos, quat, q[3 * leg:3 * leg + 3])).max() < 1e-13
            h = 1e-6
            for k in range(3):
                dq = np.zeros(3); dq[k] = h
                fd = (np_fk(m, leg, pos, quat, q[3 * leg:3 * leg + 3] + dq) - np_fk(m, leg, pos, quat, q[3 * leg:3 * leg + 3] - dq)) / (2 * h)
                assert np.abs(J[:, k] - fd).max() < 1e-8


def test_generic_axes():
    """the joint axes are parameters (the reference's URDF is external): a tilted, non-default axis set"""
    m, _ = O.leg_defaults()
    mp, _ = product_defaults()
    ax = np.array([[0.8, 0.6, 0.0], [0.0, 0.6, 0.8], [1 / 3, 2 / 3, 2 / 3]])
    for l in range(2):
        for k in range(3):
            for i in range(3):
                m.axis[9 * l + 3 * k + i] = ax[k, i]; mp.axis[9 * l + 3 * k + i] = ax[k, i]
    rng = np.random.default_rng(5)
    for _ in range(20):
        pos, quat, q = rand_state(rng)
        feet, jac = E.leg_fk(mp, pos, quat, q)
        for leg in (0, 1):
            p, J = O.leg_fk(m, leg, pos, quat, q[3 * leg:3 * leg + 3])
            assert np.abs(p - np_fk(m, leg, pos, quat, q[3 * leg:3 * leg + 3])).max() < 1e-13
            assert np.abs(feet[leg] - p).max() < 1e-13 and np.abs(jac[leg] - J).max() < 1e-13


def test_product_fk_and_torque_match_oracle():
    mo, _ = O.leg_defaults()
    mp, _ = product_defaults()
    rng = np.random.default_rng(1)
    for _ in range(100):
        pos, quat, q = rand_state(rng)
        feet, jac = E.leg_fk(mp, pos, quat, q)
        for leg in (0, 1):
            p, J = O.leg_fk(mo, leg, pos, quat, q[3 * leg:3 * leg + 3])
            assert np.abs(feet[leg] - p).max() < 1e-13 and np.abs(jac[leg] - J).max() < 1e-13
        u0 = rng.uniform(-50, 150, 6)
        u0[3 * rng.integers(0, 2):][:3] = 0.0 if rng.random() < 0.5 else u0[:3]
        tau = E.grf_to_torque(mp, quat, q, u0)
        assert np.abs(tau - O.grf_to_torque(mo, quat, q, u0)).max() < 1e-11
        # virtual work: tau . dq = -f . dp
        dq = rng.normal(size=6) * 1e-6
        f2, _ = E.leg_fk(mp, pos, quat, q + dq)
        work = -(u0.reshape(2, 3) * (f2 - feet)).sum()
        assert abs(tau @ dq - work) < 1e-9


def test_swing_step_matches_oracle():
    """gait -> foot placement -> swing profile -> IK -> cmd.q: product math vs oracle, iteration for iteration"""
    mo, po = O.leg_defaults()
    mp, pp = product_defaults()
    rng = np.random.default_rng(2)
    seen = set()
    for n in range(300):
        pos, quat, q = rand_state(rng)
        q = q * 0.5 + np.array([0.0, 0.4, -0.8, 0.0, 0.4, -0.8])     # bent knees: away from the straight-leg singularity
        des_v = rng.uniform(-1.5, 1.5, 3)
        it = int(rng.integers(0, 10_000_000)) if n % 3 else int(rng.integers(0, 2000))
        q_cmd = rng.normal(size=6)
        a = E.swing_step(mp, pp, pos, quat, q, des_v, it, q_cmd)
        b = O.swing_step(mo, po, pos, quat, q, des_v, it, q_cmd)
        assert a["leg"] == b["leg"] and a["ik_iters"] == b["ik_iters"]
        seen.add((a["leg"], a["ik_iters"] == 10))
        assert np.abs(a["feet"] - b["feet"]).max() < 1e-13
        assert np.abs(a["next_foot"] - b["next_foot"]).max() < 1e-13
        # near-singular Jacobians (damping is only 1e-6, as in the reference) amplify rounding: relative tolerance
        scale = max(1.0, np.abs(b["q_cmd"]).max()) ** 2
        assert np.abs(a["q_cmd"] - b["q_cmd"]).max() < 1e-9 * scale and abs(a["ik_err"] - b["ik_err"]) < 1e-9 * scale
        st = 3 * (1 - a["leg"])
        assert np.array_equal(a["q_cmd"][st:st + 3], q_cmd[st:st + 3])           # the stance leg's targets are untouched
        # the swing profile: height gait_height * sin(pi s), s = elapsed swing fraction
        l, r, ph, rem = O.calculate_gait(it)
        s = (0.5 - rem) / 0.5
        assert abs(a["next_foot"][2] - np.float32(0.1) * np.sin(np.pi * s)) < 1e-12
    assert {(0, True), (1, True)} <= seen or {(0, False), (1, False)} <= seen


def test_ik_converges_when_iterated():
    """DT = 0.1 moves 10 % of the damped step per iteration (reference constants); iterating the step long enough
    reaches the target to the reference tolerance 1e-3"""
    mo, po = O.leg_defaults()
    po.ik_max_iter = 200
    rng = np.random.default_rng(3)
    for _ in range(20):
        pos, quat, q = rand_state(rng)
        q3 = np.array([0.05, 0.5, -1.0]) + rng.normal(size=3) * 0.05
        target = O.leg_fk(mo, 1, pos, quat, q3 + rng.normal(size=3) * 0.15, want_jac=False)
        qs = q3.copy()
        err = C.c_double()
        its = O.lib().orc_leg_ik(C.byref(mo), C.byref(po), 1, O._p(pos), O._p(np.ascontiguousarray(quat)), O._p(target), O._p(qs), C.byref(err))
        assert its < 200 and err.value < 1e-3
        assert np.linalg.norm(O.leg_fk(mo, 1, pos, quat, qs, want_jac=False) - target) < 1.2e-3


def test_foot_placement_clamp_and_offsets():
    _, po = O.leg_defaults()
    fin = np.zeros(3)
    pos = np.array([1.0, 2.0, 0.8]); v = np.array([3.0, -3.0, 0.0])     # 0.5 * 0.5 * 3 = 0.75 -> clamped to 0.3
    O.lib().orc_foot_placement(C.byref(po), O._p(pos), O._p(v), C.c_double(0.2), 1, O._p(fin))
    assert np.allclose(fin[:2], [1.0 + 0.6 + 0.3 + po.foot_offset_left[0], 2.0 - 0.6 - 0.3 + po.foot_offset_left[1]], atol=1e-15)
    O.lib().orc_foot_placement(C.byref(po), O._p(pos), O._p(v), C.c_double(0.2), 0, O._p(fin))
    assert np.allclose(fin[:2], [1.9 + po.foot_offset_right[0], 1.1 + po.foot_offset_right[1]], atol=1e-15)


# ---- the reference's 6-D IK task as written (include/pinocchio_kinematics.h:61-149) -----------------------------------
def _exp6(xi):
    """SE(3) exponential by scipy's dense matrix exponential of the 4x4 twist (independent of oracle and product)."""
    from scipy.linalg import expm
    v, w = xi[:3], xi[3:]
    X = np.zeros((4, 4))
    X[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]); X[:3, 3] = v
    T = expm(X)
    return T[:3, :3], T[:3, 3]


def test_log6_against_scipy_matrix_logarithm():
    """log6 of Pinocchio = the se(3) matrix logarithm: [v; w] with logm(T) = [[w]x, v; 0, 0]"""
    from scipy.linalg import logm
    rng = np.random.default_rng(5)
    for scale in (1e-9, 1e-3, 0.3, 1.5, 3.0):
        for _ in range(6):
            xi = rng.standard_normal(6); xi[3:] *= scale / np.linalg.norm(xi[3:])
            R, t = _exp6(xi)
            T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
            Lg = np.real(logm(T))
            want = np.array([Lg[0, 3], Lg[1, 3], Lg[2, 3], Lg[2, 1], Lg[0, 2], Lg[1, 0]])
            tol = 1e-9 if scale < 3.0 else 1e-7
            assert np.abs(O.se3_log(R, t) - want).max() < tol
            assert np.abs(E.se3_log(R, t) - want).max() < tol
            assert np.abs(E.se3_log(R, t) - xi).max() < tol          # log6(exp6(xi)) = xi below pi


def test_jlog6_is_the_derivative_of_log6():
    """Jlog6(M) xi = d/de log6(M exp6(e xi)) (Pinocchio's convention): central differences through scipy's expm; the
    oracle forms it as the inverse of the right-Jacobian power series, the product by Pinocchio's closed form."""
    rng = np.random.default_rng(6)
    for scale in (1e-6, 0.2, 1.0, 2.5):
        xi0 = rng.standard_normal(6); xi0[3:] *= scale / np.linalg.norm(xi0[3:])
        R, t = _exp6(xi0)
        Jo, Je = O.se3_jlog(R, t), E.se3_jlog(R, t)
        assert np.abs(Jo - Je).max() < 1e-9 * max(1.0, np.abs(Jo).max())
        h = 1e-6
        for k in range(6):
            d = np.zeros(6); d[k] = h
            Rp, tp = _exp6(d); Rm, tm = _exp6(-d)
            fp = O.se3_log(R @ Rp, R @ tp + t); fm = O.se3_log(R @ Rm, R @ tm + t)
            assert np.abs((fp - fm) / (2 * h) - Jo[:, k]).max() < 2e-6


def test_reference_literal_ik_product_vs_oracle():
    """ik_mode = 1: the 6-D log6 task.  Same iterates from the product source and the oracle; the task asks a point foot
    for the identity orientation, so it does NOT converge to the position target in general (that is the reference's
    behaviour as written) -- checked: the error norm it reports includes the orientation part."""
    mo, po = O.leg_defaults(); mp, pp = product_defaults()
    po.ik_mode = 1; pp.ik_mode = 1
    rng = np.random.default_rng(8)
    for trial in range(40):
        pos, quat, q = rand_state(rng)
        if trial % 4 == 0:
            pos = np.zeros(3); quat = np.array([0, 0, 0, 1.0])     # the reference's fixed-base model: base at the origin
        leg = trial % 2
        p0 = O.leg_fk(mo, leg, pos, quat, q[3 * leg:3 * leg + 3], want_jac=False)
        target = p0 + rng.uniform(-0.05, 0.05, 3)
        for iters in (1, 3, 10):
            po.ik_max_iter = iters; pp.ik_max_iter = iters
            qo, eo, io = O.leg_ik(mo, po, leg, pos, quat, target, q[3 * leg:3 * leg + 3])
            qe, ee, ie = E.leg_ik(mp, pp, leg, pos, quat, target, q[3 * leg:3 * leg + 3])
            assert io == ie
            assert np.abs(qo - qe).max() < 1e-8 and abs(eo - ee) < 1e-8 * max(1.0, eo)
        # one step moves the joints along -J' (JJ' + damp)^-1 err, scaled by DT = 0.1: small but non-zero
        po.ik_max_iter = 1
        q1, e1, _ = O.leg_ik(mo, po, leg, pos, quat, target, q[3 * leg:3 * leg + 3])
        assert 0.0 < np.abs(q1 - q[3 * leg:3 * leg + 3]).max() < 1.0
        # the reported error is the 6-D norm: at least the orientation error of the foot frame w.r.t. the identity
        Rf = (Rotation.from_quat(quat / np.linalg.norm(quat)) *
              Rotation.from_rotvec(np.array(mo.axis).reshape(2, 3, 3)[leg][0] * q[3 * leg]) *
              Rotation.from_rotvec(np.array(mo.axis).reshape(2, 3, 3)[leg][1] * q[3 * leg + 1]) *
              Rotation.from_rotvec(np.array(mo.axis).reshape(2, 3, 3)[leg][2] * q[3 * leg + 2]))
        assert e1 >= Rf.magnitude() - 1e-9
    # the swing step honours the switch too
    pos, quat, q = rand_state(rng)
    po.ik_max_iter = 10; pp.ik_max_iter = 10
    r_o = O.swing_step(mo, po, pos, quat, q, [0.4, 0.1, 0.0], 1234, q)
    r_e = E.swing_step(mp, pp, pos, quat, q, [0.4, 0.1, 0.0], 1234, q)
    assert r_o["ik_iters"] == r_e["ik_iters"] and np.abs(r_o["q_cmd"] - r_e["q_cmd"]).max() < 1e-8
