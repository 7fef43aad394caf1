"""numpy-facing plumbing for the generic condensed-MPC C ABI (the reference QPSolver path).
Host arrays in and out, column-major like Eigen; the work runs on the device."""
import ctypes as C

import numpy as np

from . import _capi


def _F(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class LtiContext:
    def __init__(self, device=0):
        self.lib = _capi.lib()
        h = C.c_void_p()
        _capi.check(self.lib.mpc_b200_lti_create(int(device), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpc_b200_lti_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise _capi.MpcB200Error(rc, self.lib.mpc_b200_strerror(rc).decode() + " (" +
                                     self.lib.mpc_b200_lti_last_error(self.h).decode() + ")")

    def launch_count(self):
        return int(self.lib.mpc_b200_lti_launch_count(self.h))

    def discretize(self, Ac, Bc, Ts):
        Ac, Bc = _F(Ac), _F(Bc)
        NX, NU = Bc.shape
        Ad = np.zeros((NX, NX), order="F"); Bd = np.zeros((NX, NU), order="F")
        self._check(self.lib.mpc_b200_lti_discretize(self.h, 1, NX, NU, float(Ts), _p(Ac), _p(Bc), _p(Ad), _p(Bd)))
        return Ad, Bd

    def build_qp(self, Ad, Bd, Q, R, P, x_min, x_max, u_min, u_max, N, xi0, xi_ref):
        """xi0 (NX,) or (B,NX); xi_ref (NX,N+1) or (B,NX,N+1). Returns dict of per-instance arrays (B squeezed if 1)."""
        Ad, Bd, Q, R, P = _F(Ad), _F(Bd), _F(Q), _F(R), _F(P)
        NX, NU = Bd.shape; p = NX * (N + 1); n = NU * N; mi = 2 * NX * N; me = NX * N
        xi0 = np.asarray(xi0, dtype=np.float64); xi_ref = np.asarray(xi_ref, dtype=np.float64)
        single = xi0.ndim == 1
        if single:
            xi0 = xi0[None]; xi_ref = xi_ref[None]
        B = xi0.shape[0]
        xr = np.ascontiguousarray(np.stack([x.reshape(-1, order="F") for x in xi_ref]))
        xi0 = np.ascontiguousarray(xi0)
        x_min, x_max = _F(x_min), _F(x_max)
        raw = dict(H=np.zeros((B, n * n)), f=np.zeros((B, n)), A_eq=np.zeros((B, me * n)), b_eq=np.zeros((B, me)),
                   lb=np.zeros((B, n)), ub=np.zeros((B, n)), A_ineq=np.zeros((B, mi * n)), lbA_ineq=np.zeros((B, mi)),
                   ubA_ineq=np.zeros((B, mi)), A_aug=np.zeros((B, p * NX)), B_aug=np.zeros((B, p * n)))
        self._check(self.lib.mpc_b200_lti_build_qp(
            self.h, B, NX, NU, N, _p(Ad), _p(Bd), _p(Q), _p(R), _p(P), _p(x_min), _p(x_max), float(u_min), float(u_max),
            _p(xi0), _p(xr), _p(raw["H"]), _p(raw["f"]), _p(raw["A_eq"]), _p(raw["b_eq"]), _p(raw["lb"]), _p(raw["ub"]),
            _p(raw["A_ineq"]), _p(raw["lbA_ineq"]), _p(raw["ubA_ineq"]), _p(raw["A_aug"]), _p(raw["B_aug"])))
        shapes = dict(H=(n, n), A_eq=(me, n), A_ineq=(mi, n), A_aug=(p, NX), B_aug=(p, n))
        out = {}
        for k, v in raw.items():
            if k in shapes:
                v = np.stack([x.reshape(shapes[k], order="F") for x in v])
            out[k] = v[0] if single else v
        return out

    def qp_solve(self, H, f, A, lbA, ubA, lb, ub):
        H = _F(H); n = H.shape[0]; f = _F(f); lb = _F(lb); ub = _F(ub)
        m = 0 if A is None else A.shape[0]
        A_ = _F(A) if m else None; lbA_ = _F(lbA) if m else None; ubA_ = _F(ubA) if m else None
        U = np.zeros(n); st = np.zeros(1, np.int32); it = np.zeros(1, np.int32)
        self._check(self.lib.mpc_b200_qp_solve_dense(self.h, 1, n, m, _p(H), _p(f), _p(A_), _p(lb), _p(ub), _p(lbA_), _p(ubA_),
                                                     _p(U), _p(st), _p(it)))
        return U, int(st[0]), int(it[0])

    def update_state(self, Ad, Bd, xi, u):
        Ad, Bd = _F(Ad), _F(Bd); NX, NU = Bd.shape
        xi = np.array(xi, dtype=np.float64); u = np.array(u, dtype=np.float64)
        self._check(self.lib.mpc_b200_lti_update_state(self.h, 1, NX, NU, _p(Ad), _p(Bd), _p(xi), _p(u)))
        return xi
