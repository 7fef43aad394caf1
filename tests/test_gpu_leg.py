"""GPU tier: the leg kinematics kernels (csrc/leg_b200.cu) through the C ABI against the oracle
(oracle/leg_oracle.c) on the same seeded inputs, plus the full per-step pipeline
swing step -> force MPC -> joint torques for a batch.  FP64 tolerances: FK / Jacobians 1e-12, IK iterates 1e-9
relative (ill-conditioned near singular poses, as in the reference: damping 1e-6), gait/leg indices bit-exact."""
import ctypes as C

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import oracle_lib as O
from mpc_limx_control_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    return torch


def batch(seed, B):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-1, 1, (B, 3)) + np.array([0, 0, 0.8])
    quat = Rotation.from_euler("xyz", rng.uniform([-0.3, -0.3, -np.pi], [0.3, 0.3, np.pi], (B, 3))).as_quat()
    q = rng.uniform(-0.3, 0.3, (B, 6)) + np.array([0.0, 0.4, -0.8, 0.0, 0.4, -0.8])
    dv = rng.uniform(-1.5, 1.5, (B, 3))
    it = rng.integers(0, 10_000_000, B).astype(np.int32)
    it[: B // 4] = rng.integers(0, 2000, B // 4)
    # every other robot is in a physically plausible walking state (feet near the ground under the hips, small tilt,
    # moderate speed): the IK steps stay small there and the comparison with the oracle is tight
    h = slice(0, B, 2)
    n = len(range(B)[h])
    pos[h, 2] = 0.655 + rng.uniform(-0.01, 0.01, n)
    quat[h] = Rotation.from_euler("xyz", rng.uniform([-0.03, -0.03, -np.pi], [0.03, 0.03, np.pi], (n, 3))).as_quat()
    q[h] = rng.uniform(-0.05, 0.05, (n, 6)) + np.array([0.0, 0.4, -0.8, 0.0, 0.4, -0.8])
    yaw = Rotation.from_quat(quat[h]).as_euler("zyx")[:, 0]
    sp = rng.uniform(-0.3, 0.3, n)
    dv[h] = np.stack([sp * np.cos(yaw), sp * np.sin(yaw), np.zeros(n)], 1)
    return pos, quat, q, dv, it


def test_fk_and_jacobian(torch_cuda):
    torch = torch_cuda
    from mpc_limx_control_b200.leg import LegKinematics
    mo, _ = O.leg_defaults()
    for B in (1, 127, 128, 1000):       # ragged CTA tails
        pos, quat, q, _, _ = batch(B, B)
        lk = LegKinematics()
        feet, jac = lk.fk(torch.from_numpy(pos).cuda(), torch.from_numpy(quat).cuda(), torch.from_numpy(q).cuda(), want_jac=True)
        feet = feet.cpu().numpy(); jac = jac.cpu().numpy()
        f2 = lk.fk(torch.from_numpy(pos).cuda(), torch.from_numpy(quat).cuda(), torch.from_numpy(q).cuda()).cpu().numpy()
        assert np.array_equal(f2, feet)
        for b in range(0, B, max(1, B // 64)):
            for leg in (0, 1):
                p, J = O.leg_fk(mo, leg, pos[b], quat[b], q[b, 3 * leg:3 * leg + 3])
                assert np.abs(feet[b, leg] - p).max() < 1e-12 and np.abs(jac[b, leg] - J).max() < 1e-12


def test_swing_step_vs_oracle(torch_cuda):
    torch = torch_cuda
    from mpc_limx_control_b200.leg import LegKinematics
    mo, po = O.leg_defaults()
    B = 777
    pos, quat, q, dv, it = batch(11, B)
    q_cmd0 = np.random.default_rng(4).normal(size=(B, 6))
    lk = LegKinematics()
    q_cmd = torch.from_numpy(q_cmd0.copy()).cuda()
    out = lk.swing_step(torch.from_numpy(pos).cuda(), torch.from_numpy(quat).cuda(), torch.from_numpy(q).cuda(),
                        torch.from_numpy(dv).cuda(), torch.from_numpy(it).cuda(), q_cmd)
    torch.cuda.synchronize()
    qc = q_cmd.cpu().numpy(); out = {k: v.cpu().numpy() for k, v in out.items()}
    legs = set()
    n_tight = 0
    for b in range(B):
        r = O.swing_step(mo, po, pos[b], quat[b], q[b], dv[b], int(it[b]), q_cmd0[b])
        assert out["swing_leg"][b] == r["leg"] and out["ik_iters"][b] == r["ik_iters"]      # bit-exact indices
        legs.add(r["leg"])
        assert np.abs(out["feet"][b] - r["feet"]).max() < 1e-12
        assert np.abs(out["next_foot"][b] - r["next_foot"]).max() < 1e-12
        # well-conditioned iterations agree to rounding; where the damped step blew up near a singular pose (the
        # reference's damping is only 1e-6) ten iterations amplify FMA-contraction differences: loose relative bound
        sl = 3 * r["leg"]
        step = np.abs(r["q_cmd"][sl:sl + 3] - q[b, sl:sl + 3]).max()
        tol = 1e-10 if step < 0.3 else 1e-4 * max(1.0, step) ** 2
        n_tight += step < 0.3
        assert np.abs(qc[b] - r["q_cmd"]).max() < tol and abs(out["ik_err"][b] - r["ik_err"]) < tol
        st = 3 * (1 - r["leg"])
        assert np.array_equal(qc[b, st:st + 3], q_cmd0[b, st:st + 3])
    assert legs == {0, 1} and n_tight > B // 4


def test_torque_and_host_variants(torch_cuda):
    torch = torch_cuda
    from mpc_limx_control_b200 import _capi
    from mpc_limx_control_b200.leg import LegKinematics
    mo, po = O.leg_defaults()
    B = 300
    pos, quat, q, dv, it = batch(21, B)
    u0 = np.random.default_rng(9).uniform(-40, 160, (B, 6))
    u0[::2, :3] = 0.0; u0[1::2, 3:] = 0.0
    lk = LegKinematics()
    tau = lk.grf_to_torque(torch.from_numpy(quat).cuda(), torch.from_numpy(q).cuda(), torch.from_numpy(u0).cuda()).cpu().numpy()
    for b in range(B):
        assert np.abs(tau[b] - O.grf_to_torque(mo, quat[b], q[b], u0[b])).max() < 1e-10
    assert np.all(tau[::2, :3] == 0.0) and np.all(tau[1::2, 3:] == 0.0)        # swing leg: zero force -> zero torque
    # host-buffer variants agree bit for bit with the device entry points
    L = _capi.lib()
    p = lambda a: C.c_void_p(a.ctypes.data)
    tau_h = np.zeros((B, 6))
    assert L.mpc_b200_grf_to_torque_host(0, C.byref(lk.model), B, p(quat), p(q), p(u0), p(tau_h)) == 0
    assert np.array_equal(tau_h, tau)
    feet_h = np.zeros((B, 2, 3)); jac_h = np.zeros((B, 2, 3, 3))
    assert L.mpc_b200_leg_fk_host(0, C.byref(lk.model), B, p(pos), p(quat), p(q), p(feet_h), p(jac_h)) == 0
    feet_d, jac_d = lk.fk(torch.from_numpy(pos).cuda(), torch.from_numpy(quat).cuda(), torch.from_numpy(q).cuda(), want_jac=True)
    assert np.array_equal(feet_h, feet_d.cpu().numpy()) and np.array_equal(jac_h, jac_d.cpu().numpy())
    qc_h = np.zeros((B, 6)); nf = np.zeros((B, 3)); leg = np.zeros(B, np.int32); err = np.zeros(B); its = np.zeros(B, np.int32)
    assert L.mpc_b200_swing_step_host(0, C.byref(lk.model), C.byref(lk.swing), B, p(pos), p(quat), p(q), p(dv), p(it), p(qc_h), p(feet_h),
                                      p(nf), p(leg), p(err), p(its)) == 0
    qc_d = torch.zeros((B, 6), dtype=torch.float64, device="cuda")
    out = lk.swing_step(torch.from_numpy(pos).cuda(), torch.from_numpy(quat).cuda(), torch.from_numpy(q).cuda(),
                        torch.from_numpy(dv).cuda(), torch.from_numpy(it).cuda(), qc_d)
    assert np.array_equal(qc_h, qc_d.cpu().numpy()) and np.array_equal(nf, out["next_foot"].cpu().numpy())
    assert np.array_equal(leg, out["swing_leg"].cpu().numpy())
    assert L.mpc_b200_leg_fk_host(0, None, B, p(pos), p(quat), p(q), p(feet_h), None) == _capi.EINVAL


def test_full_control_step_pipeline(torch_cuda):
    """One control step for a batch of robots, everything on the device: swing step (FK feet + IK targets) ->
    force MPC with those feet and the gait clock -> tau = -J'f.  Checks each stage against the oracle chain and
    that the stance leg's torque balances the solved force (virtual work) while the swing leg gets targets."""
    torch = torch_cuda
    from mpc_limx_control_b200.engine import Engine
    from mpc_limx_control_b200.leg import LegKinematics
    N, B, Ts = 10, 64, 0.005
    d = synth.tron1_batch(77, B, N, Ts)
    rng = np.random.default_rng(8)
    pos = d["x0"][:, 3:6].copy()
    quat = Rotation.from_euler("xyz", d["x0"][:, 0:3]).as_quat()
    q = rng.uniform(-0.05, 0.05, (B, 6)) + np.array([0.0, 0.35, -0.7, 0.0, 0.35, -0.7])
    dv = np.stack([d["velocity_x"], np.zeros(B), np.zeros(B)], 1)
    it = d["iter"].copy()
    lk = LegKinematics()
    eng = Engine(horizon=N, max_batch=B, device=0, Ts=Ts)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    q_cmd = t(q.copy())
    sw = lk.swing_step(t(pos), t(quat), t(q), t(dv), t(it), q_cmd)
    feet = sw["feet"].clone()
    feet[:, :, 2] = 0.0      # stance feet are on the ground plane in the MPC model (synthetic joint angles do not touch down exactly)
    xr = eng.reference(t(d["x0"]), t(d["omega_yaw"]), t(d["velocity_x"]))
    F, st, _ = eng.solve(t(d["x0"]), xr, feet.contiguous(), it=t(it))
    u0 = F[:, 0, :].contiguous()
    tau = lk.grf_to_torque(t(quat), t(q), u0)
    torch.cuda.synchronize()
    assert int((st != 0).sum()) == 0
    u0n = u0.cpu().numpy(); taun = tau.cpu().numpy(); legn = sw["swing_leg"].cpu().numpy()
    mo, _ = O.leg_defaults()
    for b in range(B):
        sl = 3 * int(legn[b])
        assert np.all(u0n[b, sl:sl + 3] == 0.0) and np.all(taun[b, sl:sl + 3] == 0.0)       # swing foot carries no force
        assert u0n[b, 3 * (1 - int(legn[b])) + 2] > 0.0                                      # the stance foot pushes up
        assert np.abs(taun[b] - O.grf_to_torque(mo, quat[b], q[b], u0n[b])).max() < 1e-10
    eng.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_inverse_kinematics_both_tasks(torch_cuda, mode):
    """mpc_b200_leg_ik_device: the position task and the reference's 6-D log6 task as written
    (include/pinocchio_kinematics.h:61-149), every robot of the batch against the oracle; host variant identical."""
    torch = torch_cuda
    from mpc_limx_control_b200 import _capi
    from mpc_limx_control_b200.leg import LegKinematics, default_swing
    mo, po = O.leg_defaults()
    po.ik_mode = mode
    sw = default_swing(); sw.ik_mode = mode
    B = 300
    pos, quat, q, _, _ = batch(77 + mode, B)
    rng = np.random.default_rng(3)
    leg = rng.integers(0, 2, B).astype(np.int32)
    lk = LegKinematics(swing=sw)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    feet = lk.fk(t(pos), t(quat), t(q)).cpu().numpy()
    target = feet[np.arange(B), leg] + rng.uniform(-0.04, 0.04, (B, 3))
    q_out, err, its = lk.ik(t(pos), t(quat), t(leg), t(target), t(q))
    q_out, err, its = q_out.cpu().numpy(), err.cpu().numpy(), its.cpu().numpy()
    for b in range(B):
        l = int(leg[b])
        qo, eo, io = O.leg_ik(mo, po, l, pos[b], quat[b], target[b], q[b, 3 * l:3 * l + 3])
        assert io == its[b]
        assert np.abs(q_out[b, 3 * l:3 * l + 3] - qo).max() < 1e-8
        assert np.array_equal(q_out[b, 3 * (1 - l):3 * (1 - l) + 3], q[b, 3 * (1 - l):3 * (1 - l) + 3])   # other leg untouched
        assert abs(err[b] - eo) < 1e-8 * max(1.0, eo)
    if mode == 0:
        # ten steps of DT = 0.1 shrink a position error by about 0.9^10 = 0.35 (the reference's constants): it has moved
        d0 = np.linalg.norm(target - feet[np.arange(B), leg], axis=1)
        assert (err < d0).all()
    # host variant: bit-identical
    L = _capi.lib()
    qh = np.zeros((B, 6)); eh = np.zeros(B); ih = np.zeros(B, np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    pos_c, quat_c, q_c, tg_c = (np.ascontiguousarray(a) for a in (pos, quat, q, target))
    assert L.mpc_b200_leg_ik_host(0, C.byref(lk.model), C.byref(sw), B, p(pos_c), p(quat_c), p(leg), p(tg_c), p(q_c), p(qh), p(eh), p(ih)) == 0
    assert np.array_equal(qh, q_out) and np.array_equal(ih, its)
