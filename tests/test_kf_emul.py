"""CPU tier for the base-state Kalman filter (reference include/stateEstimator.h:184-337, a file
in none of the reference's build targets): the dense oracle (oracle/kf_oracle.c) against a numpy witness, and the
product's structured kernel mathematics (csrc/kf_core.cuh, host build) against the oracle over filter runs with
contact switching.  Parity with the reference itself is unpinned for this block (see oracle/kf_oracle.c)."""
import ctypes as C

import numpy as np
from scipy.spatial.transform import Rotation

import emul_lib as E
import oracle_lib as O
from mpc_limx_control_b200 import _capi


def product_defaults():
    k, m = _capi.KfParams(), _capi.LegModel()
    L = _capi.lib()
    assert L.mpc_b200_kf_default_params(C.byref(k)) == 0 and L.mpc_b200_leg_default_model(C.byref(m)) == 0
    return k, m


def np_kf(k, m, dt, quat, gyro, accel, q, dq, contact, xhat, P):
    """numpy witness of include/stateEstimator.h:217-310 (dense matrices, numpy.linalg.solve)"""
    a = np.eye(12); a[0:3, 3:6] = dt * np.eye(3)
    b = np.zeros((12, 3)); b[0:3] = 0.5 * dt * dt * np.eye(3); b[3:6] = dt * np.eye(3)
    c = np.zeros((14, 12))
    c1 = np.hstack([np.eye(3), np.zeros((3, 3))]); c2 = np.hstack([np.zeros((3, 3)), np.eye(3)])
    c[0:3, 0:6] = c1; c[3:6, 0:6] = c1; c[0:6, 6:12] = -np.eye(6); c[6:9, 0:6] = c2; c[9:12, 0:6] = c2; c[12, 8] = 1; c[13, 11] = 1
    qd = np.concatenate([np.full(3, dt / np.float32(20) * k.imu_process_noise_position),
                         np.full(3, dt * float(np.float32(9.81)) / 20.0 * k.imu_process_noise_velocity),
                         np.full(6, dt * k.foot_process_noise_position)])
    rd = np.concatenate([np.full(6, k.foot_sensor_noise_position), np.full(6, k.foot_sensor_noise_velocity), np.full(2, k.foot_height_sensor_noise)])
    Rb = Rotation.from_quat(np.asarray(quat) / np.linalg.norm(quat)).as_matrix()
    ps, vs = np.zeros(6), np.zeros(6)
    wg = Rb @ gyro
    for i in range(2):
        p, J = O.leg_fk(m, i, [0, 0, 0], quat, q[3 * i:3 * i + 3])
        v = np.cross(wg, p) + J @ dq[3 * i:3 * i + 3]
        f = 1.0 if contact[i] else k.high_suspect_number
        qd[6 + 3 * i:9 + 3 * i] *= f; rd[3 * i:3 * i + 3] *= f; rd[6 + 3 * i:9 + 3 * i] *= f; rd[12 + i] *= f
        ps[3 * i:3 * i + 3] = -p; ps[3 * i + 2] += k.foot_radius; vs[3 * i:3 * i + 3] = -v
    acc = (Rb.T if k.accel_transpose else Rb) @ accel + np.array([0, 0, -9.81])
    y = np.concatenate([ps, vs, [0, 0]])
    x = a @ xhat + b @ acc
    pm = a @ P @ a.T + np.diag(qd)
    s = c @ pm @ c.T + np.diag(rd)
    x = x + pm @ c.T @ np.linalg.solve(s, y - c @ x)
    Pn = (np.eye(12) - pm @ c.T @ np.linalg.solve(s, c)) @ pm
    Pn = (Pn + Pn.T) / 2
    if np.linalg.det(Pn[0:2, 0:2]) > 1e-6:
        Pn[0:2, 2:] = 0; Pn[2:, 0:2] = 0; Pn[0:2, 0:2] /= 10.0
    return x, Pn


def rand_inputs(rng):
    quat = Rotation.from_euler("xyz", rng.uniform([-0.2, -0.2, -np.pi], [0.2, 0.2, np.pi])).as_quat()
    return dict(quat=quat, gyro=rng.normal(size=3) * 0.3, accel=rng.normal(size=3) * 0.5 + np.array([0, 0, 9.81]),
                q=rng.uniform(-0.2, 0.2, 6) + np.array([0, 0.4, -0.8, 0, 0.4, -0.8]), dq=rng.normal(size=6) * 0.5)


def test_oracle_matches_numpy_witness():
    ko = O.kf_defaults(); mo, _ = O.leg_defaults()
    rng = np.random.default_rng(0)
    for n in range(40):
        u = rand_inputs(rng)
        A = rng.normal(size=(12, 12)); P = A @ A.T * (10.0 ** rng.uniform(-3, 2)) + 1e-3 * np.eye(12)
        xhat = rng.normal(size=12); contact = rng.integers(0, 2, 2)
        dt = float(rng.choice([0.001, 0.002, 0.005]))
        ko.accel_transpose = n % 2
        x1, P1, od = O.kf_update(ko, mo, dt, u["quat"], u["gyro"], u["accel"], u["q"], u["dq"], contact, xhat, P)
        x2, P2 = np_kf(ko, mo, dt, u["quat"], u["gyro"], u["accel"], u["q"], u["dq"], contact, xhat, P)
        sc = max(1.0, np.abs(P).max())
        assert np.abs(x1 - x2).max() < 1e-9 * max(1.0, np.abs(x2).max()) and np.abs(P1 - P2).max() < 1e-9 * sc
        Rb = Rotation.from_quat(u["quat"]).as_matrix()
        assert np.abs(od[0:3] - x1[0:3]).max() == 0 and np.abs(od[7:10] - Rb.T @ x1[3:6]).max() < 1e-13
        assert np.array_equal(od[3:7], u["quat"]) and np.array_equal(od[10:13], u["gyro"])


def test_product_math_matches_oracle_over_filter_runs():
    """200-step runs from the reference's initial condition (xHat = 0, P = 100 I) with a walking contact pattern:
    the structured/Cholesky product path and the dense/LU oracle stay together"""
    ko = O.kf_defaults(); mo, _ = O.leg_defaults()
    kp, mp = product_defaults()
    for f, _ in O.KfParams._fields_:
        assert getattr(ko, f) == getattr(kp, f), f
    rng = np.random.default_rng(1)
    for run in range(4):
        xo = np.zeros(12); Po = 100.0 * np.eye(12); xp = xo.copy(); Pp = Po.copy()
        dt = [0.001, 0.002, 0.005, 0.001][run]
        for step in range(200):
            u = rand_inputs(rng)
            contact = [1, 1] if run == 0 else [int((step // 25) % 2 == 0), int((step // 25) % 2 == 1)]
            xo, Po, odo = O.kf_update(ko, mo, dt, u["quat"], u["gyro"], u["accel"], u["q"], u["dq"], contact, xo, Po)
            xp, Pp, odp = E.kf_update(kp, mp, dt, u["quat"], u["gyro"], u["accel"], u["q"], u["dq"], contact, xp, Pp)
            assert np.abs(xp - xo).max() < 1e-9 * max(1.0, np.abs(xo).max()), (run, step)
            assert np.abs(Pp - Po).max() < 1e-9 * max(1.0, np.abs(Po).max()), (run, step)
            assert np.abs(odp - odo).max() < 1e-9 * max(1.0, np.abs(odo).max())
            assert np.array_equal(Pp, Pp.T) and np.linalg.eigvalsh(Pp).min() > -1e-9
            xp, Pp = xo.copy(), Po.copy()      # re-synchronise: compare one step at a time (the filter is contractive anyway)


def test_standing_robot_estimates_height_and_zero_velocity():
    """physics sanity: a motionless robot standing on both feet (IMU reads +g, accel_transpose irrelevant for a level
    base) converges to base height = -(foot z relative to base) + foot radius and zero velocity"""
    kp, mp = product_defaults()
    q = np.array([0, 0.4, -0.8, 0, 0.4, -0.8])
    quat = np.array([0, 0, 0, 1.0])
    feet, _ = E.leg_fk(mp, [0, 0, 0], quat, q)
    x = np.zeros(12); P = 100.0 * np.eye(12)
    for _ in range(400):
        x, P, od = E.kf_update(kp, mp, 0.002, quat, np.zeros(3), np.array([0, 0, 9.81]), q, np.zeros(6), [1, 1], x, P)
    assert abs(x[2] - (-feet[0, 2] + kp.foot_radius)) < 2e-3 and np.abs(x[3:6]).max() < 1e-3
    assert abs(x[8]) < 2e-3 and abs(x[11]) < 2e-3          # both feet on the ground plane (height measurements)
    assert np.abs((x[6:9] - x[0:3]) - (feet[0] - [0, 0, kp.foot_radius])).max() < 2e-3
