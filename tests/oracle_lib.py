"""ctypes binding of the CPU oracle (oracle/libmpc_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libmpc_oracle.so")
INFTY = 1.0e20


class GaitParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("mpc_step", C.c_int), ("swing_time", C.c_float), ("stance_time", C.c_float)]


class Tron1Params(C.Structure):
    _fields_ = [("Ts", C.c_double), ("mass", C.c_double), ("inertia", C.c_double * 9), ("q", C.c_double * 13),
                ("r", C.c_double), ("p_scale", C.c_double), ("mu", C.c_double), ("f_max", C.c_double),
                ("ltv", C.c_int), ("per_step_feet", C.c_int)]


def build(force=False):
    srcs = [os.path.join(_ROOT, "oracle", f) for f in ("mpc_oracle.c", "leg_oracle.c", "kf_oracle.c", "mpc_oracle.h", "leg_oracle.h")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_tron1_natural_residual.restype = C.c_double
    return _lib


def F(a):
    """column-major float64 copy"""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def expm(A):
    A = F(A); n = A.shape[0]; E = np.zeros((n, n), order="F")
    lib().orc_expm(n, _p(A), _p(E))
    return E


def matpow(A, k):
    A = F(A); n = A.shape[0]; P = np.zeros((n, n), order="F")
    lib().orc_matpow(n, _p(A), int(k), _p(P))
    return P


def discretize(Ac, Bc, Ts):
    Ac, Bc = F(Ac), F(Bc); NX, NU = Bc.shape
    Ad = np.zeros((NX, NX), order="F"); Bd = np.zeros((NX, NU), order="F")
    lib().orc_discretize(NX, NU, _p(Ac), _p(Bc), C.c_double(Ts), _p(Ad), _p(Bd))
    return Ad, Bd


def build_qp_params(Ad, Bd, Q, R, P, x_min, x_max, u_min, u_max, N, xi0, xi_ref):
    Ad, Bd, Q, R, P = F(Ad), F(Bd), F(Q), F(R), F(P)
    NX, NU = Bd.shape; p = NX * (N + 1); n = NU * N
    x_min, x_max, xi0, xi_ref = F(x_min), F(x_max), F(xi0), F(xi_ref)
    o = dict(A_aug=np.zeros((p, NX), order="F"), B_aug=np.zeros((p, n), order="F"), H=np.zeros((n, n), order="F"),
             f=np.zeros(n), A_eq=np.zeros((NX * N, n), order="F"), b_eq=np.zeros(NX * N), lb=np.zeros(n), ub=np.zeros(n),
             A_ineq=np.zeros((2 * NX * N, n), order="F"), lbA_ineq=np.zeros(2 * NX * N), ubA_ineq=np.zeros(2 * NX * N))
    lib().orc_build_qp_params(NX, NU, N, _p(Ad), _p(Bd), _p(Q), _p(R), _p(P), _p(x_min), _p(x_max),
                              C.c_double(u_min), C.c_double(u_max), _p(xi0), _p(xi_ref),
                              _p(o["A_aug"]), _p(o["B_aug"]), _p(o["H"]), _p(o["f"]), _p(o["A_eq"]), _p(o["b_eq"]),
                              _p(o["lb"]), _p(o["ub"]), _p(o["A_ineq"]), _p(o["lbA_ineq"]), _p(o["ubA_ineq"]))
    return o


def update_state(Ad, Bd, xi, u):
    Ad, Bd = F(Ad), F(Bd); NX, NU = Bd.shape
    xi = np.array(xi, dtype=np.float64); u = np.array(u, dtype=np.float64)
    lib().orc_update_state(NX, NU, _p(Ad), _p(Bd), _p(xi), _p(u))
    return xi


def qp_solve(H, f, A, lbA, ubA, lb, ub):
    H = F(H); n = H.shape[0]
    f = F(f); lb = F(lb); ub = F(ub)
    if A is None or A.shape[0] == 0:
        mA = 0; A_ = lbA_ = ubA_ = None
    else:
        A_ = F(A); mA = A_.shape[0]; lbA_ = F(lbA); ubA_ = F(ubA)
    u = np.zeros(n); yb = np.zeros(n); yr = np.zeros(max(mA, 1)); it = C.c_int(0)
    st = lib().orc_qp_solve(n, _p(H), _p(f), mA, _p(A_), _p(lbA_), _p(ubA_), _p(lb), _p(ub),
                            _p(u), _p(yb), _p(yr), C.byref(it))
    res = np.zeros(4)
    lib().orc_kkt_residual(n, _p(H), _p(f), mA, _p(A_), _p(lbA_), _p(ubA_), _p(lb), _p(ub), _p(u), _p(yb), _p(yr), _p(res))
    return u, dict(status=st, iters=it.value, y_bnd=yb, y_row=yr[:mA], kkt=res)


def gait_defaults():
    g = GaitParams(); lib().orc_gait_defaults(C.byref(g)); return g


def calculate_gait(it, g=None):
    g = g or gait_defaults()
    l = C.c_int(); r = C.c_int(); ph = C.c_double(); rem = C.c_double()
    lib().orc_calculate_gait(C.byref(g), int(it), C.byref(l), C.byref(r), C.byref(ph), C.byref(rem))
    return l.value, r.value, ph.value, rem.value


def contact_schedule(it, N, g=None):
    g = g or gait_defaults()
    c = np.zeros((N, 2), np.uint8)
    lib().orc_contact_schedule(C.byref(g), int(it), N, c.ctypes.data_as(C.POINTER(C.c_uint8)))
    return c


def tron1_defaults(**kw):
    p = Tron1Params(); lib().orc_tron1_defaults(C.byref(p))
    for k, v in kw.items():
        if k in ("inertia", "q"):
            arr = np.asarray(v, float).reshape(-1)
            for i, x in enumerate(arr):
                getattr(p, k)[i] = x
        else:
            setattr(p, k, v)
    return p


def tron1_model(p, yaw, pos, feet):
    Ac = np.zeros((13, 13), order="F"); Bc = np.zeros((13, 6), order="F")
    pos = np.ascontiguousarray(pos, dtype=np.float64); feet = np.ascontiguousarray(feet, dtype=np.float64)
    lib().orc_tron1_model(C.byref(p), C.c_double(yaw), _p(pos), _p(feet), _p(Ac), _p(Bc))
    return Ac, Bc


def tron1_model_literal(p, pos, foot):
    Ac = np.zeros((13, 13), order="F"); Bc = np.zeros((13, 3), order="F")
    pos = np.ascontiguousarray(pos, dtype=np.float64); foot = np.ascontiguousarray(foot, dtype=np.float64)
    lib().orc_tron1_model_literal(C.byref(p), _p(pos), _p(foot), _p(Ac), _p(Bc))
    return Ac, Bc


def tron1_reference(x0, N, Ts, omega_yaw=0.1, velocity_x=0.5):
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    xr = np.zeros((13, N + 1), order="F")
    lib().orc_tron1_reference(_p(x0), N, C.c_double(Ts), C.c_double(omega_yaw), C.c_double(velocity_x), _p(xr))
    return xr


def tron1_condense(p, N, x0, x_ref, feet, want_pred=True):
    """x_ref: (N+1,13) step-major (== column-major 13 x (N+1))."""
    n = 6 * N; pp = 13 * (N + 1)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    x_ref = np.ascontiguousarray(x_ref, dtype=np.float64); feet = np.ascontiguousarray(feet, dtype=np.float64)
    A_aug = np.zeros((pp, 13), order="F") if want_pred else None
    B_aug = np.zeros((pp, n), order="F") if want_pred else None
    H = np.zeros((n, n), order="F"); f = np.zeros(n)
    lib().orc_tron1_condense(C.byref(p), N, _p(x0), _p(x_ref), _p(feet), _p(A_aug), _p(B_aug), _p(H), _p(f))
    return dict(A_aug=A_aug, B_aug=B_aug, H=H, f=f)


def tron1_constraints(p, N, contact):
    n = 6 * N; m = 8 * N
    contact = np.ascontiguousarray(contact, dtype=np.uint8)
    A = np.zeros((m, n), order="F"); lbA = np.zeros(m); ubA = np.zeros(m); lb = np.zeros(n); ub = np.zeros(n)
    lib().orc_tron1_constraints(C.byref(p), N, contact.ctypes.data_as(C.POINTER(C.c_uint8)), _p(A), _p(lbA), _p(ubA), _p(lb), _p(ub))
    return A, lbA, ubA, lb, ub


def tron1_natural_residual(p, N, H, f, contact, u):
    H = F(H); f = F(f); u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
    contact = np.ascontiguousarray(contact, dtype=np.uint8)
    return lib().orc_tron1_natural_residual(C.byref(p), N, _p(H), _p(f), contact.ctypes.data_as(C.POINTER(C.c_uint8)), _p(u))


def tron1_solve_batch(p, N, x0, x_ref, feet, contact, nthreads=1):
    x0 = np.ascontiguousarray(x0, dtype=np.float64); B = x0.shape[0]
    x_ref = np.ascontiguousarray(x_ref, dtype=np.float64); feet = np.ascontiguousarray(feet, dtype=np.float64)
    contact = np.ascontiguousarray(contact, dtype=np.uint8)
    forces = np.zeros((B, N, 6)); status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32)
    lib().orc_tron1_solve_batch(C.byref(p), N, B, _p(x0), _p(x_ref), _p(feet), contact.ctypes.data_as(C.POINTER(C.c_uint8)),
                                _p(forces), status.ctypes.data_as(C.POINTER(C.c_int32)),
                                iters.ctypes.data_as(C.POINTER(C.c_int32)), int(nthreads))
    return forces, status, iters


def tron1_rollout(p, N, steps, x, omega_yaw, velocity_x, iter0, off_l, off_r, g=None):
    g = g or gait_defaults()
    x = np.array(x, dtype=np.float64)
    off_l = np.ascontiguousarray(off_l, dtype=np.float64); off_r = np.ascontiguousarray(off_r, dtype=np.float64)
    U = np.zeros((steps, 6))
    bad = lib().orc_tron1_rollout(C.byref(p), C.byref(g), N, int(steps), _p(x), C.c_double(omega_yaw), C.c_double(velocity_x),
                                  int(iter0), _p(off_l), _p(off_r), _p(U))
    return x, U, bad


# ---- leg kinematics oracle (oracle/leg_oracle.c) -----------------------------------------------------
class LegModel(C.Structure):
    _fields_ = [("offset", C.c_double * 30), ("axis", C.c_double * 18)]


class SwingParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("swing_time", C.c_float), ("stance_time", C.c_float), ("gait_height", C.c_float),
                ("p_rel_max", C.c_double), ("foot_offset_left", C.c_double * 3), ("foot_offset_right", C.c_double * 3),
                ("ik_tol", C.c_double), ("ik_dt", C.c_double), ("ik_damp", C.c_double), ("ik_max_iter", C.c_int32),
                ("ik_mode", C.c_int32)]


def leg_defaults():
    m, p = LegModel(), SwingParams()
    lib().orc_leg_defaults(C.byref(m), C.byref(p))
    return m, p


def _v(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def leg_fk(m, leg, pos, quat, q3, want_jac=True):
    pos, quat, q3 = _v(pos), _v(quat), _v(q3)
    p = np.zeros(3); J = np.zeros((3, 3))
    lib().orc_leg_fk(C.byref(m), int(leg), _p(pos), _p(quat), _p(q3), _p(p), _p(J) if want_jac else None)
    return (p, J) if want_jac else p


def swing_step(m, sp, pos, quat, q6, des_v, it, q_cmd):
    pos, quat, q6, des_v = _v(pos), _v(quat), _v(q6), _v(des_v)
    q_cmd = _v(q_cmd).copy(); feet = np.zeros(6); nxt = np.zeros(3); err = C.c_double(); its = C.c_int()
    leg = lib().orc_swing_step(C.byref(m), C.byref(sp), _p(pos), _p(quat), _p(q6), _p(des_v), int(it), _p(q_cmd), _p(feet), _p(nxt),
                               C.byref(err), C.byref(its))
    return dict(leg=leg, q_cmd=q_cmd, feet=feet.reshape(2, 3), next_foot=nxt, ik_err=err.value, ik_iters=its.value)


def leg_ik(m, sp, leg, pos, quat, target, q3):
    """position task (ik_mode 0) or the reference's 6-D task (ik_mode 1); returns (q3, err, iters)"""
    pos, quat, target = _v(pos), _v(quat), _v(target)
    q = _v(q3).copy(); err = C.c_double()
    fn = lib().orc_leg_ik6 if sp.ik_mode == 1 else lib().orc_leg_ik
    its = fn(C.byref(m), C.byref(sp), int(leg), _p(pos), _p(quat), _p(target), _p(q), C.byref(err))
    return q, err.value, its


def se3_log(R, t):
    R, t = _v(R), _v(t); xi = np.zeros(6)
    lib().orc_se3_log(_p(R), _p(t), _p(xi))
    return xi


def se3_jlog(R, t):
    R, t = _v(R), _v(t); J = np.zeros((6, 6))
    lib().orc_se3_jlog(_p(R), _p(t), _p(J))
    return J


def grf_to_torque(m, quat, q6, u0):
    quat, q6, u0 = _v(quat), _v(q6), _v(u0)
    tau = np.zeros(6)
    lib().orc_grf_to_torque(C.byref(m), _p(quat), _p(q6), _p(u0), _p(tau))
    return tau


# ---- Kalman filter oracle (oracle/kf_oracle.c) ----------------------------------------------------------
class KfParams(C.Structure):
    _fields_ = [("foot_radius", C.c_double), ("imu_process_noise_position", C.c_double), ("imu_process_noise_velocity", C.c_double),
                ("foot_process_noise_position", C.c_double), ("foot_sensor_noise_position", C.c_double),
                ("foot_sensor_noise_velocity", C.c_double), ("foot_height_sensor_noise", C.c_double),
                ("high_suspect_number", C.c_double), ("accel_transpose", C.c_int32)]


def kf_defaults():
    k = KfParams(); lib().orc_kf_defaults(C.byref(k)); return k


def kf_update(k, m, dt, quat, gyro, accel, q, dq, contact, xhat, P):
    """one update; returns (xhat, P, odom) as new arrays"""
    quat, gyro, accel, q, dq = _v(quat), _v(gyro), _v(accel), _v(q), _v(dq)
    c = np.ascontiguousarray(contact, np.uint8)
    x = _v(xhat).copy(); Pn = _v(P).reshape(144).copy(); od = np.zeros(13)
    lib().orc_kf_update(C.byref(k), C.byref(m), C.c_double(dt), _p(quat), _p(gyro), _p(accel), _p(q), _p(dq),
                        c.ctypes.data_as(C.POINTER(C.c_uint8)), _p(x), _p(Pn), _p(od))
    return x, Pn.reshape(12, 12), od
