"""Host-side plumbing over the C ABI: torch owns device/pinned memory and streams, the engine
computes.  Mirrors the batch entry point of BASELINE.json's north_star (set state, reference,
gait/contact schedule and foot positions; get optimal ground-reaction forces).

No CPU fallback: constructing an Engine without the built CUDA library or without a B200-class
device raises."""
import ctypes as C

import numpy as np
import torch

from . import _capi


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


class Engine:
    """One engine per GPU (single caller).  horizon in {10, 20, 50} (rollout: 10, 20)."""

    def __init__(self, horizon=10, max_batch=4096, device=0, **param_overrides):
        self.lib = _capi.lib()
        self.N = int(horizon)
        self.max_batch = int(max_batch)
        self.device = int(device)
        self.params = _capi.default_params(**param_overrides)
        self.per_step_feet = bool(self.params.per_step_feet and self.params.ltv)
        h = C.c_void_p()
        rc = self.lib.mpc_b200_create(C.byref(self.params), self.N, self.max_batch, self.device, C.byref(h))
        _capi.check(rc)
        self.h = h
        self.tdev = torch.device("cuda", self.device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpc_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    def _check_dev(self, t, dtype, numel, name):
        if t.device != self.tdev or t.dtype != dtype or not t.is_contiguous() or t.numel() != numel:
            raise ValueError(f"{name}: expected contiguous {dtype} tensor of {numel} elements on {self.tdev}")

    def launch_count(self):
        return int(self.lib.mpc_b200_launch_count(self.h))

    HOST_AUTO, HOST_STAGED, HOST_ZEROCOPY = 0, 1, 2

    def set_host_mode(self, mode):
        """Data path of the host-buffer calls: HOST_AUTO (zero-copy when every buffer is pinned, else staged
        copies), HOST_STAGED, HOST_ZEROCOPY (error instead of falling back)."""
        _capi.check(self.lib.mpc_b200_set_host_mode(self.h, int(mode)), self.h)

    def last_host_path(self):
        """1 if the last host-buffer call ran zero-copy, 0 if it staged its copies."""
        return int(self.lib.mpc_b200_last_host_path(self.h))

    # -- device-resident entry points ------------------------------------------------------------
    def contact_schedule(self, it):
        """it: int32 [B] on device -> uint8 [B,N,2] (MPC::calculateGait over the horizon)."""
        B = it.numel()
        self._check_dev(it, torch.int32, B, "iter")
        out = torch.empty((B, self.N, 2), dtype=torch.uint8, device=self.tdev)
        _capi.check(self.lib.mpc_b200_contact_schedule_device(self.h, B, _ptr(it), _ptr(out), self._stream()), self.h)
        return out

    def solve(self, x0, x_ref, feet, contact=None, it=None, forces=None, status=None, iters=None):
        """Device tensors in, device tensors out (launched on torch's current stream).
        x0 [B,13], x_ref [B,N+1,13], feet [B,2,3] (or [B,N,2,3]), contact uint8 [B,N,2] or it int32 [B]."""
        B, N = x0.shape[0], self.N
        self._check_dev(x0, torch.float64, B * 13, "x0")
        self._check_dev(x_ref, torch.float64, B * 13 * (N + 1), "x_ref")
        self._check_dev(feet, torch.float64, B * (6 * N if self.per_step_feet else 6), "feet")
        if contact is not None:
            self._check_dev(contact, torch.uint8, B * 2 * N, "contact")
        if it is not None:
            self._check_dev(it, torch.int32, B, "iter")
        if forces is None:
            forces = torch.empty((B, N, 6), dtype=torch.float64, device=self.tdev)
        if status is None:
            status = torch.empty((B,), dtype=torch.int32, device=self.tdev)
        if iters is None:
            iters = torch.empty((B,), dtype=torch.int32, device=self.tdev)
        rc = self.lib.mpc_b200_tron1_solve_device(self.h, B, _ptr(x0), _ptr(x_ref), _ptr(feet), _ptr(contact), _ptr(it),
                                                  _ptr(forces), _ptr(status), _ptr(iters), self._stream())
        _capi.check(rc, self.h)
        return forces, status, iters

    def join(self):
        """Make torch's current stream wait for every pipelined solve issued so far (mpc_b200_join)."""
        _capi.check(self.lib.mpc_b200_join(self.h, self._stream()), self.h)

    def bind_solve(self, x0, x_ref, feet, contact=None, it=None, forces=None, status=None, iters=None, pipelined=False):
        """Validate the tensors once and return a zero-overhead callable that issues exactly one
        mpc_b200_tron1_solve_device call on torch's current stream (for hot loops: the Python-side
        checks and pointer marshalling of solve() cost more than a 4096-instance batch takes on the GPU).
        The tensors must stay alive and unmoved while the callable is used."""
        B, N = x0.shape[0], self.N
        self._check_dev(x0, torch.float64, B * 13, "x0")
        self._check_dev(x_ref, torch.float64, B * 13 * (N + 1), "x_ref")
        self._check_dev(feet, torch.float64, B * (6 * N if self.per_step_feet else 6), "feet")
        if contact is not None:
            self._check_dev(contact, torch.uint8, B * 2 * N, "contact")
        if it is not None:
            self._check_dev(it, torch.int32, B, "iter")
        for t, dt, n, nm in ((forces, torch.float64, B * 6 * N, "forces"), (status, torch.int32, B, "status"), (iters, torch.int32, B, "iters")):
            if t is not None:
                self._check_dev(t, dt, n, nm)
        args = (self.h, B, _ptr(x0), _ptr(x_ref), _ptr(feet), _ptr(contact), _ptr(it), _ptr(forces), _ptr(status), _ptr(iters),
                self._stream())
        # pipelined=True: mpc_b200_tron1_solve_device_pipelined -- independent batches overlap on engine-owned streams; call
        # join() before anything on the current stream reads the results
        fn = self.lib.mpc_b200_tron1_solve_device_pipelined if pipelined else self.lib.mpc_b200_tron1_solve_device
        keep = (x0, x_ref, feet, contact, it, forces, status, iters)

        def call(_fn=fn, _args=args, _keep=keep):
            rc = _fn(*_args)
            if rc:
                _capi.check(rc, self.h)
        return call

    def condense(self, x0, x_ref, feet, want_pred=True):
        """Parity dump: H [B,n,n], f [B,n], A_aug [B,13(N+1),13], B_aug [B,13(N+1),n] as numpy,
        column-major per instance (returned arrays are indexed [b][row][col])."""
        B, N = x0.shape[0], self.N
        n, p = 6 * N, 13 * (N + 1)
        H = torch.empty((B, n, n), dtype=torch.float64, device=self.tdev)
        f = torch.empty((B, n), dtype=torch.float64, device=self.tdev)
        A = torch.empty((B, 13, p), dtype=torch.float64, device=self.tdev) if want_pred else None
        Bm = torch.empty((B, n, p), dtype=torch.float64, device=self.tdev) if want_pred else None
        rc = self.lib.mpc_b200_tron1_condense_device(self.h, B, _ptr(x0), _ptr(x_ref), _ptr(feet), _ptr(H), _ptr(f),
                                                     _ptr(A), _ptr(Bm), self._stream())
        _capi.check(rc, self.h)
        torch.cuda.synchronize(self.tdev)
        out = dict(H=H.cpu().numpy().transpose(0, 2, 1), f=f.cpu().numpy())
        if want_pred:
            out["A_aug"] = A.cpu().numpy().transpose(0, 2, 1)
            out["B_aug"] = Bm.cpu().numpy().transpose(0, 2, 1)
        return out

    def reference(self, x0, omega_yaw, velocity_x):
        """mpcQP's reference generator for a batch (device tensors): x_ref [B,N+1,13]."""
        B = x0.shape[0]
        self._check_dev(x0, torch.float64, B * 13, "x0")
        self._check_dev(omega_yaw, torch.float64, B, "omega_yaw")
        self._check_dev(velocity_x, torch.float64, B, "velocity_x")
        xr = torch.empty((B, self.N + 1, 13), dtype=torch.float64, device=self.tdev)
        _capi.check(self.lib.mpc_b200_tron1_reference_device(self.h, B, _ptr(x0), _ptr(omega_yaw), _ptr(velocity_x), _ptr(xr),
                                                             self._stream()), self.h)
        return xr

    def rollout(self, x, omega_yaw, velocity_x, it0, steps, want_traj=False):
        """Closed-loop rollout; x [B,13] is updated in place. Returns (u_traj or None, uncertified[B], iters[B])."""
        B = x.shape[0]
        self._check_dev(x, torch.float64, B * 13, "x")
        self._check_dev(omega_yaw, torch.float64, B, "omega_yaw")
        self._check_dev(velocity_x, torch.float64, B, "velocity_x")
        self._check_dev(it0, torch.int32, B, "iter0")
        traj = torch.empty((B, steps, 6), dtype=torch.float64, device=self.tdev) if want_traj else None
        bad = torch.zeros((B,), dtype=torch.int32, device=self.tdev)
        its = torch.zeros((B,), dtype=torch.int32, device=self.tdev)
        _capi.check(self.lib.mpc_b200_tron1_rollout_device(self.h, B, int(steps), _ptr(x), _ptr(omega_yaw), _ptr(velocity_x),
                                                           _ptr(it0), _ptr(traj), _ptr(bad), _ptr(its), self._stream()), self.h)
        return traj, bad, its

    # -- host-buffer entry point (the reference-facing call: H2D + solve + D2H inside) -------------
    def solve_host(self, x0, x_ref, feet, contact=None, it=None, forces=None, status=None, iters=None):
        """numpy arrays or CPU torch tensors (pinned recommended) in and out."""
        def as_np(a, dt):
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                a = a.numpy()
            a = np.ascontiguousarray(a, dtype=dt)
            return a
        x0 = as_np(x0, np.float64); x_ref = as_np(x_ref, np.float64); feet = as_np(feet, np.float64)
        contact = as_np(contact, np.uint8); it = as_np(it, np.int32)
        B, N = x0.shape[0], self.N
        if forces is None:
            forces = np.empty((B, N, 6))
        if status is None:
            status = np.empty(B, np.int32)
        if iters is None:
            iters = np.empty(B, np.int32)
        fo, so, io = as_np_out(forces), as_np_out(status), as_np_out(iters)
        p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        rc = self.lib.mpc_b200_tron1_solve_host(self.h, B, p(x0), p(x_ref), p(feet), p(contact), p(it), p(fo), p(so), p(io))
        _capi.check(rc, self.h)
        return forces, status, iters


def _np(a, dt):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        a = a.numpy()
    return np.ascontiguousarray(a, dtype=dt)


def control_host(eng, x0, omega_yaw, velocity_x, feet, contact=None, it=None, u0=None, status=None, iters=None):
    """Controller-shaped host call: state + velocity command + feet + gait -> first-step forces [B,6]."""
    x0 = _np(x0, np.float64); oy = _np(omega_yaw, np.float64); vx = _np(velocity_x, np.float64); feet = _np(feet, np.float64)
    contact = _np(contact, np.uint8); it = _np(it, np.int32)
    B = x0.shape[0]
    u0 = np.empty((B, 6)) if u0 is None else u0
    status = np.empty(B, np.int32) if status is None else status
    iters = np.empty(B, np.int32) if iters is None else iters
    p = lambda a: None if a is None else C.c_void_p(as_np_out(a).ctypes.data)
    rc = eng.lib.mpc_b200_tron1_control_host(eng.h, B, p(x0), p(oy), p(vx), p(feet), p(contact), p(it), p(u0), p(status), p(iters))
    _capi.check(rc, eng.h)
    return u0, status, iters


def _host_ptr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        assert a.device.type == "cpu" and a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def wait(eng):
    """Block until every batch queued with the asynchronous host entry is solved and copied back (mpc_b200_wait)."""
    _capi.check(eng.lib.mpc_b200_wait(eng.h), eng.h)


def bind_solve_host(eng, x0, x_ref, feet, contact=None, it=None, forces=None, status=None, iters=None, asynchronous=False):
    """Pre-marshalled mpc_b200_tron1_solve_host call (buffers: contiguous float64/uint8/int32 numpy arrays or
    CPU torch tensors, pinned recommended; they must stay alive and unmoved).  asynchronous=True binds
    mpc_b200_tron1_solve_host_async (pinned buffers required; results valid after wait(eng))."""
    B = x0.shape[0]
    args = (eng.h, B, _host_ptr(x0), _host_ptr(x_ref), _host_ptr(feet), _host_ptr(contact), _host_ptr(it), _host_ptr(forces),
            _host_ptr(status), _host_ptr(iters))
    fn = eng.lib.mpc_b200_tron1_solve_host_async if asynchronous else eng.lib.mpc_b200_tron1_solve_host
    keep = (x0, x_ref, feet, contact, it, forces, status, iters)

    def call(_fn=fn, _args=args, _keep=keep):
        rc = _fn(*_args)
        if rc:
            _capi.check(rc, eng.h)
    return call


def bind_control_host(eng, x0, omega_yaw, velocity_x, feet, contact=None, it=None, u0=None, status=None, iters=None, asynchronous=False):
    """Pre-marshalled mpc_b200_tron1_control_host call (see bind_solve_host); asynchronous=True binds the _async entry."""
    B = x0.shape[0]
    args = (eng.h, B, _host_ptr(x0), _host_ptr(omega_yaw), _host_ptr(velocity_x), _host_ptr(feet), _host_ptr(contact),
            _host_ptr(it), _host_ptr(u0), _host_ptr(status), _host_ptr(iters))
    fn = eng.lib.mpc_b200_tron1_control_host_async if asynchronous else eng.lib.mpc_b200_tron1_control_host
    keep = (x0, omega_yaw, velocity_x, feet, contact, it, u0, status, iters)

    def call(_fn=fn, _args=args, _keep=keep):
        rc = _fn(*_args)
        if rc:
            _capi.check(rc, eng.h)
    return call


def solve_host_multi(engines, x0, x_ref, feet, contact=None, it=None, forces=None, status=None, iters=None):
    """Single-process multi-GPU batch solve (mpc_b200_tron1_solve_host_multi): `engines` = one Engine per device, the batch
    is block-partitioned by instance, every GPU reads/writes its rows of the caller's host arrays."""
    e0 = engines[0]
    x0 = _np(x0, np.float64); x_ref = _np(x_ref, np.float64); feet = _np(feet, np.float64)
    contact = _np(contact, np.uint8); it = _np(it, np.int32)
    B, N = x0.shape[0], e0.N
    forces = np.empty((B, N, 6)) if forces is None else forces
    status = np.empty(B, np.int32) if status is None else status
    iters = np.empty(B, np.int32) if iters is None else iters
    hs = (C.c_void_p * len(engines))(*[e.h for e in engines])
    p = lambda a: None if a is None else C.c_void_p(as_np_out(a).ctypes.data)
    rc = e0.lib.mpc_b200_tron1_solve_host_multi(hs, len(engines), B, p(x0), p(x_ref), p(feet), p(contact), p(it), p(forces),
                                                p(status), p(iters))
    for e in engines:
        if rc:
            _capi.check(rc, e.h)
    return forces, status, iters


def pin_host_buffer(a):
    """Page-lock a numpy array in place (mpc_b200_pin_host_buffer) so that host calls using it run zero-copy."""
    _capi.check(_capi.lib().mpc_b200_pin_host_buffer(C.c_void_p(a.ctypes.data), a.nbytes))
    return a


def unpin_host_buffer(a):
    _capi.check(_capi.lib().mpc_b200_unpin_host_buffer(C.c_void_p(a.ctypes.data)))


def as_np_out(a):
    if isinstance(a, torch.Tensor):
        return a.numpy()
    return a


def measure_fp64_peak(device=0):
    v = C.c_double(0.0)
    _capi.check(_capi.lib().mpc_b200_measure_fp64_peak(int(device), C.byref(v)))
    return v.value
