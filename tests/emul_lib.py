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
    ox = 0.05556 - 0.077 - 0.15 + 0.145 + 0.0; oz = -0.2602 + 0.0 - 0.25981 - 0.2598 - 0.032
    for i, v in enumerate([ox, -0.105 - 0.0205 - (-0.0205), oz]): p.foot_offset_left[i] = v
    for i, v in enumerate([ox, 0.105 + 0.0205 + (-0.0205), oz]): p.foot_offset_right[i] = v
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


def solve_tiled60(p, x0, x_ref, feet, contact):
    """Horizon 10, 60-variable capacity class in the tiled 8x8 storage (the latency class of the kernel wrapper)."""
    N = 10
    x0 = np.ascontiguousarray(x0, np.float64); x_ref = np.ascontiguousarray(x_ref, np.float64)
    feet = np.ascontiguousarray(feet, np.float64); contact = np.ascontiguousarray(contact, np.uint8)
    forces = np.zeros((N, 6)); it = C.c_int(0)
    st = lib().emul_tron1_solve_tiled60(C.byref(p), x0.ctypes.data_as(_dp), x_ref.ctypes.data_as(_dp),
                                        feet.ctypes.data_as(_dp), contact.ctypes.data_as(_u8),
                                        forces.ctypes.data_as(_dp), C.byref(it))
    return forces, st, it.value


def solve_riccati(p, N, x0, x_ref, feet, contact, ext_gains=False):
    """Riccati work type (O(N) active-face solves); returns (forces, status, iters, deferred-to-dense flag)."""
    x0 = np.ascontiguousarray(x0, np.float64); x_ref = np.ascontiguousarray(x_ref, np.float64)
    feet = np.ascontiguousarray(feet, np.float64); contact = np.ascontiguousarray(contact, np.uint8)
    forces = np.zeros((N, 6)); it = C.c_int(0); df = C.c_int(0)
    st = lib().emul_tron1_solve_riccati(C.byref(p), N, int(ext_gains), x0.ctypes.data_as(_dp), x_ref.ctypes.data_as(_dp),
                                        feet.ctypes.data_as(_dp), contact.ctypes.data_as(_u8),
                                        forces.ctypes.data_as(_dp), C.byref(it), C.byref(df))
    return forces, st, it.value, df.value


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


def rollout(p, N, steps, x, omega_yaw, velocity_x, iter0):
    x = np.array(x, dtype=np.float64)
    U = np.zeros((steps, 6)); it = C.c_int(0)
    bad = lib().emul_tron1_rollout(C.byref(p), N, int(steps), x.ctypes.data_as(_dp), C.c_double(omega_yaw),
                                   C.c_double(velocity_x), int(iter0), U.ctypes.data_as(_dp), C.byref(it))
    return x, U, bad, it.value


def gait_contact(p, it, N):
    c = np.zeros((N, 2), np.uint8)
    lib().emul_gait_contact(C.byref(p), int(it), N, c.ctypes.data_as(_u8))
    return c


# ---- generic LTI path (csrc/lti_core.cuh) -----------------------------------------------------------
_lti = None


def lti():
    global _lti
    if _lti is None:
        lib()
        _lti = C.CDLL(os.path.join(_ROOT, "tests", "emul", "libemul_lti.so"))
    return _lti


def _F(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _pp(a):
    return None if a is None else a.ctypes.data_as(_dp)


def lti_discretize(Ac, Bc, Ts):
    Ac, Bc = _F(Ac), _F(Bc); NX, NU = Bc.shape
    Ad = np.zeros((NX, NX), order="F"); Bd = np.zeros((NX, NU), order="F")
    lti().emul_lti_discretize(NX, NU, C.c_double(Ts), _pp(Ac), _pp(Bc), _pp(Ad), _pp(Bd))
    return Ad, Bd


def lti_build(Ad, Bd, Q, R, P, x_min, x_max, u_min, u_max, N, xi0, xi_ref):
    Ad, Bd, Q, R, P = _F(Ad), _F(Bd), _F(Q), _F(R), _F(P); NX, NU = Bd.shape; p = NX * (N + 1); n = NU * N
    x_min, x_max, xi0, xi_ref = _F(x_min), _F(x_max), _F(xi0), _F(xi_ref)
    o = dict(H=np.zeros((n, n), order="F"), f=np.zeros(n), A_eq=np.zeros((NX * N, n), order="F"), b_eq=np.zeros(NX * N),
             lb=np.zeros(n), ub=np.zeros(n), A_ineq=np.zeros((2 * NX * N, n), order="F"), lbA_ineq=np.zeros(2 * NX * N),
             ubA_ineq=np.zeros(2 * NX * N), A_aug=np.zeros((p, NX), order="F"), B_aug=np.zeros((p, n), order="F"))
    lti().emul_lti_build(NX, NU, N, _pp(Ad), _pp(Bd), _pp(Q), _pp(R), _pp(P), _pp(x_min), _pp(x_max), C.c_double(u_min),
                         C.c_double(u_max), _pp(xi0), _pp(xi_ref), _pp(o["H"]), _pp(o["f"]), _pp(o["A_eq"]), _pp(o["b_eq"]),
                         _pp(o["lb"]), _pp(o["ub"]), _pp(o["A_ineq"]), _pp(o["lbA_ineq"]), _pp(o["ubA_ineq"]),
                         _pp(o["A_aug"]), _pp(o["B_aug"]))
    return o


def qp_dense(H, f, A, lbA, ubA, lb, ub, max_newton=30, max_admm=4000):
    H = _F(H); n = H.shape[0]; f = _F(f); lb = _F(lb); ub = _F(ub)
    m = 0 if A is None else A.shape[0]
    A_ = _F(A) if m else None; lbA_ = _F(lbA) if m else None; ubA_ = _F(ubA) if m else None
    U = np.zeros(n); it = C.c_int(0)
    st = lti().emul_qp_dense(n, m, _pp(H), _pp(f), _pp(A_), _pp(lb), _pp(ub), _pp(lbA_), _pp(ubA_), _pp(U), C.byref(it),
                             int(max_newton), int(max_admm))
    return U, st, it.value


def lti_update(Ad, Bd, xi, u):
    Ad, Bd = _F(Ad), _F(Bd); NX, NU = Bd.shape
    xi = np.array(xi, dtype=np.float64); u = np.array(u, dtype=np.float64)
    lti().emul_lti_update(NX, NU, _pp(Ad), _pp(Bd), _pp(xi), _pp(u))
    return xi


# ---- leg kinematics (csrc/leg_core.cuh) -----------------------------------------------------------------
_leg = None


def leg_lib():
    global _leg
    if _leg is None:
        subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "tests", "emul")])
        _leg = C.CDLL(os.path.join(_ROOT, "tests", "emul", "libemul_leg.so"))
    return _leg


def leg_fk(m, pos, quat, q6):
    pos = np.ascontiguousarray(pos, np.float64); quat = np.ascontiguousarray(quat, np.float64); q6 = np.ascontiguousarray(q6, np.float64)
    feet = np.zeros((2, 3)); jac = np.zeros((2, 3, 3))
    leg_lib().emul_leg_fk(C.byref(m), pos.ctypes.data_as(_dp), quat.ctypes.data_as(_dp), q6.ctypes.data_as(_dp),
                          feet.ctypes.data_as(_dp), jac.ctypes.data_as(_dp))
    return feet, jac


def swing_step(m, sp, pos, quat, q6, des_v, it, q_cmd):
    a = lambda x: np.ascontiguousarray(x, np.float64)
    pos, quat, q6, des_v = a(pos), a(quat), a(q6), a(des_v)
    q_cmd = a(q_cmd).copy(); feet = np.zeros(6); nxt = np.zeros(3); err = C.c_double(); its = C.c_int()
    leg = leg_lib().emul_swing_step(C.byref(m), C.byref(sp), pos.ctypes.data_as(_dp), quat.ctypes.data_as(_dp), q6.ctypes.data_as(_dp),
                                    des_v.ctypes.data_as(_dp), int(it), q_cmd.ctypes.data_as(_dp), feet.ctypes.data_as(_dp),
                                    nxt.ctypes.data_as(_dp), C.byref(err), C.byref(its))
    return dict(leg=leg, q_cmd=q_cmd, feet=feet.reshape(2, 3), next_foot=nxt, ik_err=err.value, ik_iters=its.value)


def leg_ik(m, sp, leg, pos, quat, target, q3):
    a = lambda x: np.ascontiguousarray(x, np.float64)
    pos, quat, target = a(pos), a(quat), a(target)
    q = a(q3).copy(); err = C.c_double()
    its = leg_lib().emul_leg_ik(C.byref(m), C.byref(sp), int(leg), pos.ctypes.data_as(_dp), quat.ctypes.data_as(_dp),
                                target.ctypes.data_as(_dp), q.ctypes.data_as(_dp), C.byref(err))
    return q, err.value, its


def se3_log(R, t):
    a = lambda x: np.ascontiguousarray(x, np.float64)
    R, t = a(R), a(t); xi = np.zeros(6)
    leg_lib().emul_se3_log(R.ctypes.data_as(_dp), t.ctypes.data_as(_dp), xi.ctypes.data_as(_dp))
    return xi


def se3_jlog(R, t):
    a = lambda x: np.ascontiguousarray(x, np.float64)
    R, t = a(R), a(t); J = np.zeros((6, 6))
    leg_lib().emul_se3_jlog(R.ctypes.data_as(_dp), t.ctypes.data_as(_dp), J.ctypes.data_as(_dp))
    return J


def grf_to_torque(m, quat, q6, u0):
    a = lambda x: np.ascontiguousarray(x, np.float64)
    quat, q6, u0 = a(quat), a(q6), a(u0)
    tau = np.zeros(6)
    leg_lib().emul_grf_to_torque(C.byref(m), quat.ctypes.data_as(_dp), q6.ctypes.data_as(_dp), u0.ctypes.data_as(_dp), tau.ctypes.data_as(_dp))
    return tau


def kf_update(k, m, dt, quat, gyro, accel, q, dq, contact, xhat, P):
    a = lambda x: np.ascontiguousarray(x, np.float64)
    quat, gyro, accel, q, dq = a(quat), a(gyro), a(accel), a(q), a(dq)
    c = np.ascontiguousarray(contact, np.uint8)
    x = a(xhat).copy(); Pn = a(P).reshape(144).copy(); od = np.zeros(13)
    leg_lib().emul_kf_update(C.byref(k), C.byref(m), C.c_double(dt), quat.ctypes.data_as(_dp), gyro.ctypes.data_as(_dp),
                             accel.ctypes.data_as(_dp), q.ctypes.data_as(_dp), dq.ctypes.data_as(_dp), c.ctypes.data_as(_u8),
                             x.ctypes.data_as(_dp), Pn.ctypes.data_as(_dp), od.ctypes.data_as(_dp))
    return x, Pn.reshape(12, 12), od
