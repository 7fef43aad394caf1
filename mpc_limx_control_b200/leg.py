"""Host-side plumbing of the leg kinematics kernels (csrc/leg_b200.cu) over the C ABI: the rest of the
reference's MPC::run around the force MPC (include/MPCController.h:183-196) for a batch of robots.
torch owns the device memory and the stream; nothing here computes."""
import ctypes as C

import torch

from . import _capi


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class LegKinematics:
    """Batched FK / swing-leg step / GRF->torque on one GPU (device tensors in, device tensors out)."""

    def __init__(self, device=0, model=None, swing=None):
        self.lib = _capi.lib()
        self.model = model or default_model()
        self.swing = swing or default_swing()
        self.tdev = torch.device("cuda", int(device))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    def _chk(self, t, dtype, numel, name):
        if t.device != self.tdev or t.dtype != dtype or not t.is_contiguous() or t.numel() != numel:
            raise ValueError(f"{name}: expected contiguous {dtype} tensor of {numel} elements on {self.tdev}")

    def fk(self, base_pos, base_quat, q, want_jac=False):
        """base_pos [B,3], base_quat [B,4] ([x,y,z,w]), q [B,6] -> feet [B,2,3] (and jac [B,2,3,3])."""
        B = q.shape[0]
        self._chk(base_pos, torch.float64, 3 * B, "base_pos"); self._chk(base_quat, torch.float64, 4 * B, "base_quat")
        self._chk(q, torch.float64, 6 * B, "q")
        feet = torch.empty((B, 2, 3), dtype=torch.float64, device=self.tdev)
        jac = torch.empty((B, 2, 3, 3), dtype=torch.float64, device=self.tdev) if want_jac else None
        with torch.cuda.device(self.tdev):
            _capi.check(self.lib.mpc_b200_leg_fk_device(C.byref(self.model), B, _ptr(base_pos), _ptr(base_quat), _ptr(q), _ptr(feet),
                                                        _ptr(jac), self._stream()))
        return (feet, jac) if want_jac else feet

    def swing_step(self, base_pos, base_quat, q, des_vel, it, q_cmd):
        """One swing-leg step per robot; q_cmd [B,6] is updated in place (swing leg entries only).
        Returns dict(feet [B,2,3], next_foot [B,3], swing_leg [B], ik_err [B], ik_iters [B])."""
        B = q.shape[0]
        for t, n, nm in ((base_pos, 3, "base_pos"), (base_quat, 4, "base_quat"), (q, 6, "q"), (des_vel, 3, "des_vel"), (q_cmd, 6, "q_cmd")):
            self._chk(t, torch.float64, n * B, nm)
        self._chk(it, torch.int32, B, "iter")
        feet = torch.empty((B, 2, 3), dtype=torch.float64, device=self.tdev)
        nxt = torch.empty((B, 3), dtype=torch.float64, device=self.tdev)
        leg = torch.empty(B, dtype=torch.int32, device=self.tdev)
        err = torch.empty(B, dtype=torch.float64, device=self.tdev)
        its = torch.empty(B, dtype=torch.int32, device=self.tdev)
        with torch.cuda.device(self.tdev):
            _capi.check(self.lib.mpc_b200_swing_step_device(C.byref(self.model), C.byref(self.swing), B, _ptr(base_pos), _ptr(base_quat),
                                                            _ptr(q), _ptr(des_vel), _ptr(it), _ptr(q_cmd), _ptr(feet), _ptr(nxt), _ptr(leg),
                                                            _ptr(err), _ptr(its), self._stream()))
        return dict(feet=feet, next_foot=nxt, swing_leg=leg, ik_err=err, ik_iters=its)

    def ik(self, base_pos, base_quat, leg, target, q_init):
        """PinocchioKinematics::inverseKinematics for a batch: leg [B] int32 (0 left, 1 right), target [B,3] world position
        of the contact point, q_init [B,6] -> (q_out [B,6], ik_err [B], ik_iters [B]).  self.swing.ik_mode selects the
        position task (0) or the reference's 6-D log6 task as written (1)."""
        B = q_init.shape[0]
        for t, n, nm in ((base_pos, 3, "base_pos"), (base_quat, 4, "base_quat"), (target, 3, "target"), (q_init, 6, "q_init")):
            self._chk(t, torch.float64, n * B, nm)
        self._chk(leg, torch.int32, B, "leg")
        q_out = torch.empty((B, 6), dtype=torch.float64, device=self.tdev)
        err = torch.empty(B, dtype=torch.float64, device=self.tdev)
        its = torch.empty(B, dtype=torch.int32, device=self.tdev)
        with torch.cuda.device(self.tdev):
            _capi.check(self.lib.mpc_b200_leg_ik_device(C.byref(self.model), C.byref(self.swing), B, _ptr(base_pos), _ptr(base_quat),
                                                        _ptr(leg), _ptr(target), _ptr(q_init), _ptr(q_out), _ptr(err), _ptr(its),
                                                        self._stream()))
        return q_out, err, its

    def grf_to_torque(self, base_quat, q, u0, tau=None):
        """tau [B,6] = -J(q)' f per leg from the first-step forces u0 [B,6]."""
        B = q.shape[0]
        self._chk(base_quat, torch.float64, 4 * B, "base_quat"); self._chk(q, torch.float64, 6 * B, "q"); self._chk(u0, torch.float64, 6 * B, "u0")
        if tau is None:
            tau = torch.empty((B, 6), dtype=torch.float64, device=self.tdev)
        with torch.cuda.device(self.tdev):
            _capi.check(self.lib.mpc_b200_grf_to_torque_device(C.byref(self.model), B, _ptr(base_quat), _ptr(q), _ptr(u0), _ptr(tau),
                                                               self._stream()))
        return tau


def default_model():
    m = _capi.LegModel()
    _capi.check(_capi.lib().mpc_b200_leg_default_model(C.byref(m)))
    return m


def default_swing():
    p = _capi.SwingParams()
    _capi.check(_capi.lib().mpc_b200_swing_default_params(C.byref(p)))
    return p


class StateEstimator:
    """Batched base-state Kalman filter (csrc/kf_b200.cu; reference include/stateEstimator.h:184-337): xhat [B,12] and
    P [B,12,12] stay on the device between updates."""

    def __init__(self, B, device=0, model=None, params=None, p0=100.0):
        self.lib = _capi.lib()
        self.B = int(B)
        self.tdev = torch.device("cuda", int(device))
        self.model = model or default_model()
        self.params = params or default_kf_params()
        self.xhat = torch.empty((self.B, 12), dtype=torch.float64, device=self.tdev)
        self.P = torch.empty((self.B, 12, 12), dtype=torch.float64, device=self.tdev)
        self.reset(p0)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    def reset(self, p0=100.0):
        with torch.cuda.device(self.tdev):
            _capi.check(self.lib.mpc_b200_kf_reset_device(self.B, float(p0), _ptr(self.xhat), _ptr(self.P), self._stream()))

    def update(self, dt, quat, gyro, accel, q, dq, contact, odom=None):
        """One filter update for every robot; returns odom [B,13] = pos, quat, body-frame velocity, angular velocity."""
        B = self.B
        for t, n, nm in ((quat, 4, "quat"), (gyro, 3, "gyro"), (accel, 3, "accel"), (q, 6, "q"), (dq, 6, "dq")):
            if t.device != self.tdev or t.dtype != torch.float64 or not t.is_contiguous() or t.numel() != n * B:
                raise ValueError(f"{nm}: expected contiguous float64 tensor of {n * B} elements on {self.tdev}")
        if contact.device != self.tdev or contact.dtype != torch.uint8 or not contact.is_contiguous() or contact.numel() != 2 * B:
            raise ValueError("contact: expected contiguous uint8 [B,2] on the device")
        if odom is None:
            odom = torch.empty((B, 13), dtype=torch.float64, device=self.tdev)
        with torch.cuda.device(self.tdev):
            _capi.check(self.lib.mpc_b200_kf_update_device(C.byref(self.params), C.byref(self.model), B, float(dt), _ptr(quat), _ptr(gyro),
                                                           _ptr(accel), _ptr(q), _ptr(dq), _ptr(contact), _ptr(self.xhat), _ptr(self.P),
                                                           _ptr(odom), self._stream()))
        return odom


def default_kf_params():
    k = _capi.KfParams()
    _capi.check(_capi.lib().mpc_b200_kf_default_params(C.byref(k)))
    return k
