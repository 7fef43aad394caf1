"""GPU tier: the batched Kalman filter kernel (csrc/kf_b200.cu) through the C ABI against the dense oracle
(oracle/kf_oracle.c) over multi-step runs with contact switching; FP64, 1e-9 relative per step."""
import ctypes as C

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    return torch


def inputs(rng, B):
    quat = Rotation.from_euler("xyz", rng.uniform([-0.2, -0.2, -np.pi], [0.2, 0.2, np.pi], (B, 3))).as_quat()
    return dict(quat=quat, gyro=rng.normal(size=(B, 3)) * 0.3, accel=rng.normal(size=(B, 3)) * 0.5 + np.array([0, 0, 9.81]),
                q=rng.uniform(-0.2, 0.2, (B, 6)) + np.array([0, 0.4, -0.8, 0, 0.4, -0.8]), dq=rng.normal(size=(B, 6)) * 0.5)


def test_filter_runs_vs_oracle(torch_cuda):
    torch = torch_cuda
    from mpc_limx_control_b200.leg import StateEstimator
    ko = O.kf_defaults(); mo, _ = O.leg_defaults()
    B, steps, dt = 37, 60, 0.002          # ragged CTA tail (8 robots per CTA)
    rng = np.random.default_rng(3)
    est = StateEstimator(B)
    xo = np.zeros((B, 12)); Po = np.tile(100.0 * np.eye(12), (B, 1, 1))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for step in range(steps):
        u = inputs(rng, B)
        contact = np.stack([(np.arange(B) + step // 10) % 2, (np.arange(B) + step // 10 + 1) % 2], 1).astype(np.uint8)
        contact[::5] = 1                  # some robots in double support
        odom = est.update(dt, t(u["quat"]), t(u["gyro"]), t(u["accel"]), t(u["q"]), t(u["dq"]), t(contact))
        torch.cuda.synchronize()
        xg = est.xhat.cpu().numpy(); Pg = est.P.cpu().numpy(); og = odom.cpu().numpy()
        for b in range(B):
            xo[b], Po[b], od = O.kf_update(ko, mo, dt, u["quat"][b], u["gyro"][b], u["accel"][b], u["q"][b], u["dq"][b], contact[b], xo[b], Po[b])
            assert np.abs(xg[b] - xo[b]).max() < 1e-9 * max(1.0, np.abs(xo[b]).max()), (step, b)
            assert np.abs(Pg[b] - Po[b]).max() < 1e-9 * max(1.0, np.abs(Po[b]).max()), (step, b)
            assert np.abs(og[b] - od).max() < 1e-9 * max(1.0, np.abs(od).max())
        assert np.array_equal(Pg, Pg.transpose(0, 2, 1))
    # FREE-RUNNING: the oracle filter and the device filter each carried their own state through all 60 updates (nothing
    # was re-synchronised), so the 1e-9 bound above is a bound on the accumulated drift, not on a single step


def test_host_variant_and_errors(torch_cuda):
    torch = torch_cuda
    from mpc_limx_control_b200 import _capi
    from mpc_limx_control_b200.leg import StateEstimator
    B, dt = 9, 0.001
    rng = np.random.default_rng(4)
    u = inputs(rng, B)
    contact = rng.integers(0, 2, (B, 2)).astype(np.uint8)
    est = StateEstimator(B, p0=3.0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    odom = est.update(dt, t(u["quat"]), t(u["gyro"]), t(u["accel"]), t(u["q"]), t(u["dq"]), t(contact))
    torch.cuda.synchronize()
    L = _capi.lib()
    p = lambda a: C.c_void_p(a.ctypes.data)
    xh = np.zeros((B, 12)); Ph = np.tile(3.0 * np.eye(12), (B, 1, 1)); oh = np.zeros((B, 13))
    assert L.mpc_b200_kf_update_host(0, C.byref(est.params), C.byref(est.model), B, dt, p(u["quat"]), p(u["gyro"]), p(u["accel"]),
                                     p(u["q"]), p(u["dq"]), p(contact), p(xh), p(Ph), p(oh)) == 0
    assert np.array_equal(xh, est.xhat.cpu().numpy()) and np.array_equal(Ph, est.P.cpu().numpy()) and np.array_equal(oh, odom.cpu().numpy())
    assert L.mpc_b200_kf_update_host(0, C.byref(est.params), C.byref(est.model), B, 0.0, p(u["quat"]), p(u["gyro"]), p(u["accel"]),
                                     p(u["q"]), p(u["dq"]), p(contact), p(xh), p(Ph), p(oh)) == _capi.EINVAL
