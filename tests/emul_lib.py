"""ctypes binding of the host emulation build of the device source (tests/emul). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import sys
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
from mpc_limx_control_b200._capi import Tron1Params  # noqa: E402  (struct layout only)

_SO = os.path.join(_ROOT, "tests", "emul", "libemul_tron1.so")
_lib = None
_dp = C.POINTER(C.c_double)
_u8 = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "tests", "emul")])
        _lib = C.CDLL(_SO)
    return _lib


def default_params(**kw):
    """Same defaults as mpc_b200_tron1_default_params (csrc/tron1_params.h)."""
    p = Tron1Params()
    p.Ts = 0.005; p.mass = 9.585
    I = [140110.479E-06, 534.939E-06, 28184.116E-06, 534.939E-06, 110641.449E-06, -27.278E-06,
         28184.116E-06, -27.278E-06, 98944.542E-06]
    q = [1, 1, 10, 100, 100, 100, 50, 50, 50, 100, 100, 100, 0.1]
    for i, v in enumerate(I): p.inertia[i] = v
    for i, v in enumerate(q): p.q[i] = v
    p.r = 0.1; p.p_scale = 20.0; p.mu = 0.5; p.f_max = 2.0 * 9.585 * 9.8
    p.ltv = 1; p.per_step_feet = 0
    p.gait_dt = 0.001; p.gait_mpc_step = 5; p.gait_swing_time = 0.5; p.gait_stance_time = 0.5
    p.max_newton = 12; p.max_admm = 2000; p.tol = 1e-9
    for k, v in kw.items():
        if k in ("inertia", "q"):
            for i, x in enumerate(np.asarray(v, float).reshape(-1)): getattr(p, k)[i] = x
        else:
            setattr(p, k, v)
    return p


def solve(p, N, x0, x_ref, feet, contact):
    x0 = np.ascontiguousarray(x0, np.float64); x_ref = np.ascontiguousarray(x_ref, np.float64)
    feet = np.ascontiguousarray(feet, np.float64); contact = np.ascontiguousarray(contact, np.uint8)
    forces = np.zeros((N, 6)); it = C.c_int(0)
    st = lib().emul_tron1_solve(C.byref(p), N, x0.ctypes.data_as(_dp), x_ref.ctypes.data_as(_dp),
                                feet.ctypes.data_as(_dp), contact.ctypes.data_as(_u8),
                                forces.ctypes.data_as(_dp), C.byref(it))
    return forces, st, it.value


def dump(p, N, x0, x_ref, feet):
    x0 = np.ascontiguousarray(x0, np.float64); x_ref = np.ascontiguousarray(x_ref, np.float64)
    feet = np.ascontiguousarray(feet, np.float64)
    n = 6 * N; pp = 13 * (N + 1)
    H = np.zeros((n, n), order="F"); f = np.zeros(n)
    A = np.zeros((pp, 13), order="F"); Bm = np.zeros((pp, n), order="F")
    rc = lib().emul_tron1_dump(C.byref(p), N, x0.ctypes.data_as(_dp), x_ref.ctypes.data_as(_dp),
                               feet.ctypes.data_as(_dp), H.ctypes.data_as(_dp), f.ctypes.data_as(_dp),
                               A.ctypes.data_as(_dp), Bm.ctypes.data_as(_dp))
    assert rc == 0
    return dict(H=H, f=f, A_aug=A, B_aug=Bm)


def gait_contact(p, it, N):
    c = np.zeros((N, 2), np.uint8)
    lib().emul_gait_contact(C.byref(p), int(it), N, c.ctypes.data_as(_u8))
    return c
