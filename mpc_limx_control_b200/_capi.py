"""ctypes binding of the C ABI (include/mpc_b200.h -> csrc/libmpc_b200.so).

Plumbing only: torch (or any allocator) owns the buffers, this module passes raw pointers.
There is no CPU fallback: if the shared library is missing or no CUDA device is usable the
calls raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPC_B200_LIB") or os.path.join(_HERE, "csrc", "libmpc_b200.so")   # override: A/B builds in tools/

OK, EINVAL, ENODEV, ECUDA, ENOMEM, ECAPACITY = 0, -1, -2, -3, -4, -5
INFTY = 1.0e20

# every symbol include/mpc_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "mpc_b200_version", "mpc_b200_strerror", "mpc_b200_device_count", "mpc_b200_measure_fp64_peak",
    "mpc_b200_tron1_default_params", "mpc_b200_create", "mpc_b200_destroy", "mpc_b200_last_error",
    "mpc_b200_launch_count", "mpc_b200_set_host_mode", "mpc_b200_last_host_path", "mpc_b200_pin_host_buffer", "mpc_b200_unpin_host_buffer", "mpc_b200_contact_schedule_device", "mpc_b200_tron1_solve_device",
    "mpc_b200_tron1_solve_host", "mpc_b200_tron1_solve_device_pipelined", "mpc_b200_join", "mpc_b200_tron1_solve_host_async", "mpc_b200_tron1_control_host_async", "mpc_b200_wait", "mpc_b200_tron1_solve_host_multi", "mpc_b200_tron1_condense_device",
    "mpc_b200_tron1_reference_device", "mpc_b200_tron1_rollout_device", "mpc_b200_tron1_control_host",
    "mpc_b200_leg_default_model", "mpc_b200_swing_default_params", "mpc_b200_leg_fk_device", "mpc_b200_swing_step_device",
    "mpc_b200_grf_to_torque_device", "mpc_b200_leg_fk_host", "mpc_b200_swing_step_host", "mpc_b200_grf_to_torque_host",
    "mpc_b200_leg_ik_device", "mpc_b200_leg_ik_host",
    "mpc_b200_kf_default_params", "mpc_b200_kf_reset_device", "mpc_b200_kf_update_device", "mpc_b200_kf_update_host",
    "mpc_b200_lti_create", "mpc_b200_lti_destroy", "mpc_b200_lti_last_error", "mpc_b200_lti_launch_count",
    "mpc_b200_lti_discretize", "mpc_b200_lti_build_qp", "mpc_b200_qp_solve_dense", "mpc_b200_lti_update_state",
]


class Tron1Params(C.Structure):
    _fields_ = [
        ("Ts", C.c_double), ("mass", C.c_double), ("inertia", C.c_double * 9), ("q", C.c_double * 13),
        ("r", C.c_double), ("p_scale", C.c_double), ("mu", C.c_double), ("f_max", C.c_double),
        ("ltv", C.c_int32), ("per_step_feet", C.c_int32),
        ("gait_dt", C.c_float), ("gait_mpc_step", C.c_int32),
        ("gait_swing_time", C.c_float), ("gait_stance_time", C.c_float),
        ("max_newton", C.c_int32), ("max_admm", C.c_int32), ("tol", C.c_double),
        ("foot_offset_left", C.c_double * 3), ("foot_offset_right", C.c_double * 3),
    ]


class LegModel(C.Structure):
    """mpc_b200_leg_model: offset[leg][link][xyz], axis[leg][joint][xyz] (flattened)."""
    _fields_ = [("offset", C.c_double * 30), ("axis", C.c_double * 18)]


class SwingParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("swing_time", C.c_float), ("stance_time", C.c_float), ("gait_height", C.c_float),
                ("p_rel_max", C.c_double), ("foot_offset_left", C.c_double * 3), ("foot_offset_right", C.c_double * 3),
                ("ik_tol", C.c_double), ("ik_dt", C.c_double), ("ik_damp", C.c_double), ("ik_max_iter", C.c_int32),
                ("ik_mode", C.c_int32)]


class KfParams(C.Structure):
    _fields_ = [("foot_radius", C.c_double), ("imu_process_noise_position", C.c_double), ("imu_process_noise_velocity", C.c_double),
                ("foot_process_noise_position", C.c_double), ("foot_sensor_noise_position", C.c_double),
                ("foot_sensor_noise_velocity", C.c_double), ("foot_height_sensor_noise", C.c_double),
                ("high_suspect_number", C.c_double), ("accel_transpose", C.c_int32)]


class MpcB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mpc_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libmpc_b200.so (built by __graft_entry__.build()). Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the engine has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, ip, dp = C.c_void_p, C.c_int, C.POINTER(C.c_double)
        L.mpc_b200_version.restype = ip
        L.mpc_b200_strerror.restype = C.c_char_p
        L.mpc_b200_strerror.argtypes = [ip]
        L.mpc_b200_device_count.restype = ip
        L.mpc_b200_measure_fp64_peak.argtypes = [ip, dp]
        L.mpc_b200_tron1_default_params.argtypes = [C.POINTER(Tron1Params)]
        L.mpc_b200_create.argtypes = [C.POINTER(Tron1Params), ip, ip, ip, C.POINTER(vp)]
        L.mpc_b200_destroy.argtypes = [vp]
        L.mpc_b200_last_error.restype = C.c_char_p
        L.mpc_b200_last_error.argtypes = [vp]
        L.mpc_b200_launch_count.restype = C.c_int64
        L.mpc_b200_launch_count.argtypes = [vp]
        L.mpc_b200_set_host_mode.argtypes = [vp, ip]
        L.mpc_b200_last_host_path.argtypes = [vp]
        L.mpc_b200_pin_host_buffer.argtypes = [vp, C.c_size_t]
        L.mpc_b200_unpin_host_buffer.argtypes = [vp]
        L.mpc_b200_contact_schedule_device.argtypes = [vp, ip, vp, vp, vp]
        L.mpc_b200_tron1_solve_device.argtypes = [vp, ip, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.mpc_b200_tron1_solve_host.argtypes = [vp, ip, vp, vp, vp, vp, vp, vp, vp, vp]
        L.mpc_b200_tron1_solve_device_pipelined.argtypes = [vp, ip, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.mpc_b200_join.argtypes = [vp, vp]
        L.mpc_b200_tron1_solve_host_async.argtypes = [vp, ip, vp, vp, vp, vp, vp, vp, vp, vp]
        L.mpc_b200_wait.argtypes = [vp]
        L.mpc_b200_tron1_control_host_async.argtypes = [vp, ip] + [vp] * 9
        L.mpc_b200_tron1_solve_host_multi.argtypes = [C.POINTER(vp), ip, ip, vp, vp, vp, vp, vp, vp, vp, vp]
        L.mpc_b200_tron1_condense_device.argtypes = [vp, ip, vp, vp, vp, vp, vp, vp, vp, vp]
        L.mpc_b200_tron1_reference_device.argtypes = [vp, ip, vp, vp, vp, vp, vp]
        L.mpc_b200_tron1_rollout_device.argtypes = [vp, ip, ip, vp, vp, vp, vp, vp, vp, vp, vp]
        L.mpc_b200_tron1_control_host.argtypes = [vp, ip] + [vp] * 9
        L.mpc_b200_leg_default_model.argtypes = [C.POINTER(LegModel)]
        L.mpc_b200_swing_default_params.argtypes = [C.POINTER(SwingParams)]
        L.mpc_b200_leg_fk_device.argtypes = [C.POINTER(LegModel), ip] + [vp] * 6
        L.mpc_b200_swing_step_device.argtypes = [C.POINTER(LegModel), C.POINTER(SwingParams), ip] + [vp] * 12
        L.mpc_b200_grf_to_torque_device.argtypes = [C.POINTER(LegModel), ip] + [vp] * 5
        L.mpc_b200_leg_ik_device.argtypes = [C.POINTER(LegModel), C.POINTER(SwingParams), ip] + [vp] * 9
        L.mpc_b200_leg_ik_host.argtypes = [ip, C.POINTER(LegModel), C.POINTER(SwingParams), ip] + [vp] * 8
        L.mpc_b200_leg_fk_host.argtypes = [ip, C.POINTER(LegModel), ip] + [vp] * 5
        L.mpc_b200_swing_step_host.argtypes = [ip, C.POINTER(LegModel), C.POINTER(SwingParams), ip] + [vp] * 11
        L.mpc_b200_grf_to_torque_host.argtypes = [ip, C.POINTER(LegModel), ip] + [vp] * 4
        L.mpc_b200_kf_default_params.argtypes = [C.POINTER(KfParams)]
        L.mpc_b200_kf_reset_device.argtypes = [ip, C.c_double, vp, vp, vp]
        L.mpc_b200_kf_update_device.argtypes = [C.POINTER(KfParams), C.POINTER(LegModel), ip, C.c_double] + [vp] * 10
        L.mpc_b200_kf_update_host.argtypes = [ip, C.POINTER(KfParams), C.POINTER(LegModel), ip, C.c_double] + [vp] * 9
        dd = C.c_double
        L.mpc_b200_lti_create.argtypes = [ip, C.POINTER(vp)]
        L.mpc_b200_lti_destroy.argtypes = [vp]
        L.mpc_b200_lti_last_error.restype = C.c_char_p
        L.mpc_b200_lti_last_error.argtypes = [vp]
        L.mpc_b200_lti_launch_count.restype = C.c_int64
        L.mpc_b200_lti_launch_count.argtypes = [vp]
        L.mpc_b200_lti_discretize.argtypes = [vp, ip, ip, ip, dd, vp, vp, vp, vp]
        L.mpc_b200_lti_build_qp.argtypes = [vp, ip, ip, ip, ip] + [vp] * 7 + [dd, dd] + [vp] * 13
        L.mpc_b200_qp_solve_dense.argtypes = [vp, ip, ip, ip] + [vp] * 10
        L.mpc_b200_lti_update_state.argtypes = [vp, ip, ip, ip, vp, vp, vp, vp]
        _lib = L
    return _lib


def default_params(**overrides):
    p = Tron1Params()
    lib().mpc_b200_tron1_default_params(C.byref(p))
    for k, v in overrides.items():
        if k in ("inertia", "q", "foot_offset_left", "foot_offset_right"):
            for i, x in enumerate(v):
                getattr(p, k)[i] = float(x)
        else:
            setattr(p, k, v)
    return p


def check(code, engine=None):
    if code != OK:
        L = lib()
        msg = L.mpc_b200_strerror(code).decode()
        if engine:
            extra = L.mpc_b200_last_error(engine).decode()
            if extra:
                msg += f" ({extra})"
        raise MpcB200Error(code, msg)
