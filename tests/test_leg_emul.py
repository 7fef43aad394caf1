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
