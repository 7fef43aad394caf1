"""Independent numpy/scipy restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY.  This module is the *second* witness used to pin the C
oracle (oracle/mpc_oracle.c): it re-derives every quantity with library routines
(scipy.linalg.expm, numpy.linalg.matrix_power, dense numpy products) instead of the
hand-written loops the oracle uses, and it solves the QP with a different algorithm
(primal-dual interior point) from the oracle's dual active-set method.

Reference citations (paths relative to /root/reference):
  discretize          src/QPSolver.cpp:21-29
  build_qp_params     src/QPSolver.cpp:31-81
  update_state        src/QPSolver.cpp:108-111
  demo scenario       src/qpSolver_test.cpp:6-50
  TRON1 constants     include/mpcQP.h:18-22,54-56,66-71,74-97
  TRON1 model         include/mpcQP.h:121-182 (literal) / BASELINE.json north_star (intended)
  gait                include/MPCController.h:61-75, include/MPCParam.h:44-49
"""
import numpy as np
import scipy.linalg as sla

INFTY = 1.0e20  # qpOASES::INFTY

# ----------------------------------------------------------------------------- generic LTI path


def discretize(Ac, Bc, Ts):
    """QPSolver::discretizeSystem (src/QPSolver.cpp:21-29)."""
    NX, NU = Bc.shape
    M = np.zeros((NX + NU, NX + NU))
    M[:NX, :NX] = Ac
    M[:NX, NX:] = Bc
    E = sla.expm(M * Ts)
    return E[:NX, :NX].copy(), E[:NX, NX:].copy()


def build_qp_params(Ad, Bd, Q, R, P, x_min, x_max, u_min, u_max, N, xi0, xi_ref):
    """QPSolver::buildQPParams (src/QPSolver.cpp:31-81). xi_ref is NX x (N+1)."""
    NX, NU = Bd.shape
    A_aug = np.zeros((NX * (N + 1), NX))
    B_aug = np.zeros((NX * (N + 1), NU * N))
    A_aug[:NX] = np.eye(NX)
    for i in range(1, N + 1):
        A_aug[i * NX:(i + 1) * NX] = Ad @ A_aug[(i - 1) * NX:i * NX]
    for i in range(1, N + 1):
        for j in range(i):
            B_aug[i * NX:(i + 1) * NX, j * NU:(j + 1) * NU] = np.linalg.matrix_power(Ad, i - j - 1) @ Bd
    Q_bar = np.zeros((NX * (N + 1), NX * (N + 1)))
    R_bar = np.zeros((NU * N, NU * N))
    for i in range(N):
        Q_bar[i * NX:(i + 1) * NX, i * NX:(i + 1) * NX] = Q
        R_bar[i * NU:(i + 1) * NU, i * NU:(i + 1) * NU] = R
    Q_bar[N * NX:, N * NX:] = P
    H = 2 * (B_aug.T @ Q_bar @ B_aug + R_bar)
    xi_ref_vec = xi_ref.reshape(-1, order="F")
    f = 2 * B_aug.T @ Q_bar @ (A_aug @ xi0 - xi_ref_vec)
    A_eq = B_aug[NX:]
    b_eq = A_aug[NX:] @ xi0
    lb = np.full(NU * N, u_min)
    ub = np.full(NU * N, u_max)
    A_ineq = np.zeros((2 * NX * N, NU * N))
    lbA = np.full(2 * NX * N, -INFTY)
    ubA = np.full(2 * NX * N, INFTY)
    for i in range(N):
        A_pred = np.linalg.matrix_power(Ad, i + 1)
        A_ineq[2 * i * NX:2 * i * NX + NX] = B_aug[(i + 1) * NX:(i + 2) * NX]
        lbA[2 * i * NX:2 * i * NX + NX] = x_min - A_pred @ xi0
        ubA[2 * i * NX:2 * i * NX + NX] = x_max - A_pred @ xi0
    return dict(A_aug=A_aug, B_aug=B_aug, H=H, f=f, A_eq=A_eq, b_eq=b_eq, lb=lb, ub=ub,
                A_ineq=A_ineq, lbA_ineq=lbA, ubA_ineq=ubA)


def demo_system():
    """src/qpSolver_test.cpp:6-24."""
    Ts, N = 0.01, 15
    Ac = np.array([[0, 1, 0, 0], [0, -0.1, 0, 0], [0, 0, 0, 1], [0, 0, 0, -0.1]], float)
    Bc = np.array([[0, 0], [5, 0], [0, 0], [0, 5]], float)
    Q = np.diag([50.0, 5, 50, 5])
    R = 0.1 * np.eye(2)
    P = 20 * Q
    x_min = np.array([-5.0, -3, -5, -3])
    return dict(Ts=Ts, N=N, Ac=Ac, Bc=Bc, Q=Q, R=R, P=P, x_min=x_min, x_max=-x_min, u_min=-8.0, u_max=8.0)


def demo_reference(k, Ts, N, radius=2.0, w=0.5):
    """src/qpSolver_test.cpp:40-50."""
    xr = np.zeros((4, N + 1))
    for i in range(N + 1):
        t = k * Ts + i * Ts
        th = w * t
        xr[0, i] = radius * np.cos(th)
        xr[2, i] = radius * np.sin(th)
        xr[1, i] = -radius * w * np.sin(th)
        xr[3, i] = radius * w * np.cos(th)
    return xr


# ----------------------------------------------------------------------------- QP (interior point)


def qp_ipm(H, f, A, lbA, ubA, lb, ub, tol=1e-11, max_iter=200):
    """min 1/2 u'Hu + f'u  s.t. lb<=u<=ub, lbA<=Au<=ubA   (src/QPSolver.cpp:83-106 problem form).

    Mehrotra predictor-corrector on G u <= h; equality rows (lo==hi) are eliminated through a
    null-space basis.  Returns (u, info)."""
    n = H.shape[0]
    rows, rhs, eqr, eqb = [], [], [], []
    I = np.eye(n)

    def add(a, lo, hi):
        if lo > -INFTY / 2 and hi < INFTY / 2 and hi - lo <= 0.0:
            eqr.append(a); eqb.append(lo); return
        if hi < INFTY / 2:
            rows.append(a); rhs.append(hi)
        if lo > -INFTY / 2:
            rows.append(-a); rhs.append(-lo)

    for i in range(n):
        add(I[i], lb[i], ub[i])
    for r in range(A.shape[0]):
        if np.any(A[r] != 0.0):
            add(A[r], lbA[r], ubA[r])
    # eliminate equalities u = u_p + Z w
    if eqr:
        E = np.array(eqr); be = np.array(eqb)
        u_p = np.linalg.lstsq(E, be, rcond=None)[0]
        Z = sla.null_space(E)
    else:
        u_p = np.zeros(n); Z = np.eye(n)
    G = np.array(rows) if rows else np.zeros((0, n))
    h = np.array(rhs) if rows else np.zeros(0)
    Hr = Z.T @ H @ Z
    fr = Z.T @ (f + H @ u_p)
    Gr = G @ Z
    hr = h - G @ u_p
    # drop rows that vanished after the elimination
    keep = np.abs(Gr).sum(axis=1) > 1e-14
    Gr, hr = Gr[keep], hr[keep]
    nr, m = Hr.shape[0], Gr.shape[0]
    if nr == 0:
        return u_p, dict(iters=0)
    w = np.linalg.solve(Hr, -fr)
    if m == 0:
        return u_p + Z @ w, dict(iters=0)
    s = np.maximum(hr - Gr @ w, 1.0)
    lam = np.ones(m)
    it = 0
    for it in range(max_iter):
        rd = Hr @ w + fr + Gr.T @ lam
        rp = Gr @ w + s - hr
        mu = s @ lam / m
        if max(np.abs(rd).max(), np.abs(rp).max()) < tol and mu < tol:
            break
        if mu < 1e-15:
            break
        D = lam / s
        K = Hr + Gr.T @ (D[:, None] * Gr)
        if not np.all(np.isfinite(K)):
            break
        try:
            cf = sla.cho_factor(K)
            ksolve = lambda b_: sla.cho_solve(cf, b_)
        except np.linalg.LinAlgError:
            ksolve = lambda b_: np.linalg.lstsq(K, b_, rcond=None)[0]

        def solve(rc):
            # rc: complementarity residual target  (s*dlam + lam*ds = -rc)
            # ds = -rp - G dw ; dlam = (-rc - lam*ds)/s
            # H dw + G' dlam = -rd -> (H + G' D G) dw = -rd + G'((rc - lam*rp)/s)
            dw = ksolve(-rd + Gr.T @ ((rc - lam * rp) / s))
            ds = -rp - Gr @ dw
            dl = (-rc - lam * ds) / s
            return dw, ds, dl

        def steplen(v, dv):
            neg = dv < 0
            return min(1.0, (-v[neg] / dv[neg]).min()) if neg.any() else 1.0

        dw_a, ds_a, dl_a = solve(s * lam)
        a_aff = min(steplen(s, ds_a), steplen(lam, dl_a))
        mu_aff = (s + a_aff * ds_a) @ (lam + a_aff * dl_a) / m
        sigma = (mu_aff / mu) ** 3
        dw, ds, dl = solve(s * lam + ds_a * dl_a - sigma * mu)
        a = 0.995 * min(steplen(s, ds), steplen(lam, dl))
        if not (np.all(np.isfinite(dw)) and np.all(np.isfinite(ds)) and np.all(np.isfinite(dl))):
            break
        w += a * dw; s = np.maximum(s + a * ds, 1e-200); lam = np.maximum(lam + a * dl, 1e-200)
    return u_p + Z @ w, dict(iters=it)


# ----------------------------------------------------------------------------- TRON1 single-rigid-body path

TRON1_MASS = 9.585  # include/mpcQP.h:18
TRON1_INERTIA = np.array([[140110.479e-06, 534.939e-06, 28184.116e-06],
                          [534.939e-06, 110641.449e-06, -27.278e-06],
                          [28184.116e-06, -27.278e-06, 98944.542e-06]])  # include/mpcQP.h:20-22
TRON1_Q = np.array([1, 1, 10, 100, 100, 100, 50, 50, 50, 100, 100, 100, 0.1], float)  # include/mpcQP.h:54
TRON1_R = 0.1   # include/mpcQP.h:55
TRON1_PSCALE = 20.0  # include/mpcQP.h:56
FOOT_OFFSET_L = np.array([0.05556 - 0.077 - 0.15 + 0.145 + 0.0,
                          -0.105 - 0.0205 - (-0.0205) + 0.0 + 0.0,
                          -0.2602 + 0.0 - 0.25981 - 0.2598 - 0.032])  # include/MPCParam.h:64-68
FOOT_OFFSET_R = np.array([0.05556 - 0.077 - 0.15 + 0.145 + 0.0,
                          0.105 + 0.0205 + (-0.0205) + 0.0 + 0.0,
                          -0.2602 + 0.0 - 0.25981 - 0.2598 - 0.032])  # include/MPCParam.h:70-73


def skew(r):
    return np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])


def rotz(psi):
    c, s = np.cos(psi), np.sin(psi)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def tron1_model(yaw, pos, feet, mass=TRON1_MASS, inertia=TRON1_INERTIA):
    """Intended single-rigid-body linearisation (north_star): 13 states, 6 forces.
    State order [rpy, p, omega, v, g] (include/mpcQP.h:66-71)."""
    Ac = np.zeros((13, 13)); Bc = np.zeros((13, 6))
    Rz = rotz(yaw)
    Ac[0:3, 6:9] = Rz.T
    Ac[3:6, 9:12] = np.eye(3)
    Ac[11, 12] = 1.0
    Iw_inv = Rz @ np.linalg.inv(inertia) @ Rz.T
    for i in range(2):
        Bc[6:9, 3 * i:3 * i + 3] = Iw_inv @ skew(feet[i] - pos)
        Bc[9:12, 3 * i:3 * i + 3] = np.eye(3) / mass
    return Ac, Bc


def tron1_model_literal(pos, foot, mass=TRON1_MASS):
    """Reference-literal model (include/mpcQP.h:139-181): one support foot, NU=3."""
    dx, dy, dz = foot - pos
    Ac = np.zeros((13, 13)); Bc = np.zeros((13, 3))
    Ac[0, 7], Ac[0, 8] = dz, dy
    Ac[1, 6], Ac[1, 8] = dz, dx
    Ac[2, 6], Ac[2, 7] = dy, dx
    Ac[3, 9] = Ac[4, 10] = Ac[5, 11] = 1.0
    Ac[11, 12] = -1.0
    Bc[9, 0] = Bc[10, 1] = Bc[11, 2] = -mass
    return Ac, Bc


def tron1_reference(x0, N, Ts, omega_yaw=0.1, velocity_x=0.5):
    """include/mpcQP.h:74-97. Returns 13 x (N+1)."""
    xr = np.zeros((13, N + 1))
    for i in range(N + 1):
        t = i * Ts
        xr[:, i] = x0
        xr[2, i] = x0[2] + t * omega_yaw
        xr[3, i] = x0[3] + t * velocity_x
        xr[9, i] = x0[9] if i == 0 else velocity_x
        xr[12, i] = -9.8
    return xr


def tron1_condense(x0, x_ref, feet, N, Ts, ltv, qdiag=TRON1_Q, rw=TRON1_R, pscale=TRON1_PSCALE,
                   mass=TRON1_MASS, inertia=TRON1_INERTIA):
    """A_aug, B_aug, H, f for the TRON1 problem.

    feet: (2,3) or (N,2,3).  ltv=0: one model at x0 (the QPSolver LTI structure,
    src/QPSolver.cpp:36-60).  ltv=1: per-step model at (k==0 ? x0 : x_ref[:,k]) with
    B_aug[i,j] = A_{i-1}...A_{j+1} B_j."""
    NX, NU = 13, 6
    feet = np.asarray(feet, float)
    if feet.ndim == 2:
        feet = np.broadcast_to(feet, (N, 2, 3))
    Ads, Bds = [], []
    for k in range(N):
        lin = x0 if (k == 0 or not ltv) else x_ref[:, k]
        fk = feet[k] if ltv else feet[0]
        Ac, Bc = tron1_model(lin[2], lin[3:6], fk, mass, inertia)
        Ad, Bd = discretize(Ac, Bc, Ts)
        Ads.append(Ad); Bds.append(Bd)
    A_aug = np.zeros((NX * (N + 1), NX)); B_aug = np.zeros((NX * (N + 1), NU * N))
    A_aug[:NX] = np.eye(NX)
    for i in range(1, N + 1):
        A_aug[i * NX:(i + 1) * NX] = Ads[i - 1] @ A_aug[(i - 1) * NX:i * NX]
        for j in range(i):
            Phi = np.eye(NX)
            for k in range(j + 1, i):
                Phi = Ads[k] @ Phi
            B_aug[i * NX:(i + 1) * NX, j * NU:(j + 1) * NU] = Phi @ Bds[j]
    Qb = np.diag(np.concatenate([np.tile(qdiag, N), pscale * qdiag]))
    Rb = rw * np.eye(NU * N)
    H = 2 * (B_aug.T @ Qb @ B_aug + Rb)
    f = 2 * B_aug.T @ Qb @ (A_aug @ x0 - x_ref.reshape(-1, order="F"))
    return dict(A_aug=A_aug, B_aug=B_aug, H=H, f=f, Ad0=Ads[0], Bd0=Bds[0])


def tron1_constraints(contact, N, mu, fmax):
    """Friction pyramid + contact bounds.
    contact: (N,2) in {0,1}.  Rows per foot-step: mu*fz -/+ fx >= 0, mu*fz -/+ fy >= 0."""
    n = 6 * N
    lb = np.zeros(n); ub = np.zeros(n)
    A = np.zeros((8 * N, n)); lbA = np.zeros(8 * N); ubA = np.full(8 * N, INFTY)
    for k in range(N):
        for i in range(2):
            b = 6 * k + 3 * i
            c = float(contact[k, i])
            lb[b:b + 2] = -INFTY * c; ub[b:b + 2] = INFTY * c
            lb[b + 2] = 0.0; ub[b + 2] = c * fmax
            r = 8 * k + 4 * i
            A[r, b + 2] = mu; A[r, b] = -1
            A[r + 1, b + 2] = mu; A[r + 1, b] = 1
            A[r + 2, b + 2] = mu; A[r + 2, b + 1] = -1
            A[r + 3, b + 2] = mu; A[r + 3, b + 1] = 1
    return A, lbA, ubA, lb, ub


def calculate_gait(it):
    """MPC::calculateGait (include/MPCController.h:61-75) with MPCParam's float members
    (include/MPCParam.h:44-49): int*float -> float, widened to double."""
    dt = np.float32(0.001)
    swing = np.float32(0.5); stance = np.float32(0.5)
    current = float(np.float32(np.float32(it) * dt))
    cycle = float(np.float32(swing + stance))
    phase = float(np.fmod(current, cycle))
    if phase < float(swing):
        return 1, 0, phase, float(swing) - phase
    return 0, 1, phase, cycle - phase


def contact_schedule(it, N, mpc_step=5):
    """contact[k][foot] = 1 when stance (leg_state == 0) at iter + k*mpcStep; foot 0 = left."""
    c = np.zeros((N, 2), np.uint8)
    for k in range(N):
        l, r, _, _ = calculate_gait(it + k * mpc_step)
        c[k, 0] = 1 - l
        c[k, 1] = 1 - r
    return c
