"""Numpy prototype of the device QP algorithm (design aid, not product, not test oracle).

ADMM on the splitting  min 1/2 u'Hu + f'u + I_C(z), u = z  with the EXACT Euclidean
projection onto the per-foot-step set C_s = {|fx|<=mu fz, |fy|<=mu fz, 0<=fz<=fmax}
(or {0} for a swing foot), followed by an active-face polish (reduced equality-constrained
solve) that is verified by the natural residual ||u - P_C(u - (Hu+f))||.
Used to choose rho / alpha / check cadence before writing the CUDA kernel.
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import npref as R
from mpc_limx_control_b200 import synth


def project(v, contact, mu, fmax):
    """v: (S,3). returns z, face codes (ax, ay in {-1,0,1}: x/y clipped to sign*mu*fz; zt in {0 free,1 at fmax,2 at 0(apex)}; 3 swing)."""
    S = v.shape[0]
    z = np.zeros_like(v)
    ax = np.zeros(S, int); ay = np.zeros(S, int); zt = np.zeros(S, int)
    for s in range(S):
        if not contact[s]:
            zt[s] = 3
            continue
        x, y, w = v[s]
        a, b = max(abs(x), abs(y)), min(abs(x), abs(y))
        if w >= a / mu:
            t = w
        else:
            t = (w + mu * a) / (1 + mu * mu)
            if t < b / mu:
                t = (w + mu * (a + b)) / (1 + 2 * mu * mu)
        if t >= fmax:
            t = fmax; zt[s] = 1
        elif t <= 0:
            t = 0.0; zt[s] = 2
        lim = mu * t
        if abs(x) > lim:
            ax[s] = 1 if x > 0 else -1
        if abs(y) > lim:
            ay[s] = 1 if y > 0 else -1
        if zt[s] == 2:
            ax[s] = ay[s] = 0
        z[s] = (np.clip(x, -lim, lim), np.clip(y, -lim, lim), t)
    return z, (ax, ay, zt)


def polish(H, f, face, mu, fmax):
    """Solve min q(u) on the affine hull of the identified face."""
    ax, ay, zt = face
    S = len(ax); n = 3 * S
    cols = []; u0 = np.zeros(n)
    for s in range(S):
        b = 3 * s
        if zt[s] >= 2:
            continue
        zdir = np.zeros(n); zdir[b + 2] = 1.0; zdir[b] = ax[s] * mu; zdir[b + 1] = ay[s] * mu
        if zt[s] == 1:
            u0 += fmax * zdir
        else:
            cols.append(zdir)
        if ax[s] == 0:
            e = np.zeros(n); e[b] = 1; cols.append(e)
        if ay[s] == 0:
            e = np.zeros(n); e[b + 1] = 1; cols.append(e)
    if not cols:
        return u0
    Z = np.array(cols).T
    Hr = Z.T @ H @ Z
    w = np.linalg.solve(Hr, -Z.T @ (f + H @ u0))
    return u0 + Z @ w


def nat_res(H, f, u, contact, mu, fmax):
    g = H @ u + f
    p, _ = project((u - g).reshape(-1, 3), contact, mu, fmax)
    return np.abs(u - p.reshape(-1)).max()


def solve(H, f, contact, mu, fmax, rho=None, alpha=1.6, max_iter=2000, check=10, tol=1e-9, verbose=False):
    n = H.shape[0]
    ev = np.linalg.eigvalsh(H)
    if rho is None:
        rho = np.sqrt(ev[0] * ev[-1])
    Kinv = np.linalg.inv(H + rho * np.eye(n))
    z = np.zeros(n); y = np.zeros(n)
    npolish = 0
    last_face = None
    for it in range(1, max_iter + 1):
        u = Kinv @ (rho * (z - y) - f)
        uh = alpha * u + (1 - alpha) * z
        zn, face = project((uh + y).reshape(-1, 3), contact, mu, fmax)
        zn = zn.reshape(-1)
        y = y + uh - zn
        z = zn
        if it % check == 0:
            key = tuple(np.concatenate(face))
            if key == last_face or True:
                up = polish(H, f, face, mu, fmax)
                npolish += 1
                r = nat_res(H, f, up, contact, mu, fmax)
                if r <= tol * max(1.0, np.abs(up).max()):
                    return up, it, npolish
            last_face = key
    return z, max_iter, npolish


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    Ts = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
    B = 60
    mu, fmax = 0.5, 2 * R.TRON1_MASS * 9.8
    d = synth.tron1_batch(1001, B, N, Ts)
    for rho_mult in (0.3, 1.0, 3.0):
        its = []; pols = []; errs = []
        for b in range(B):
            x0 = d["x0"][b]; xr = d["x_ref"][b].T
            c = R.tron1_condense(x0, xr, d["feet"][b], N, Ts, 1)
            contact = R.contact_schedule(int(d["iter"][b]), N)
            H, f = c["H"], c["f"]
            ev = np.linalg.eigvalsh(H)
            u, it, npol = solve(H, f, contact.reshape(-1), mu, fmax, rho=rho_mult * np.sqrt(ev[0] * ev[-1]))
            A, lbA, ubA, lb, ub = R.tron1_constraints(contact, N, mu, fmax)
            uo, _ = R.qp_ipm(H, f, A, lbA, ubA, lb, ub)
            its.append(it); pols.append(npol); errs.append(np.abs(u - uo).max() / max(1, np.abs(uo).max()))
        its = np.array(its)
        print(f"rho x{rho_mult}: iters mean {its.mean():.1f} p50 {np.median(its)} p90 {np.percentile(its,90)} max {its.max()}  polish mean {np.mean(pols):.1f}  err max {max(errs):.2e}")
