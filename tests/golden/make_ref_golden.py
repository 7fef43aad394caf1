"""Generates tests/golden/ref_v1.npz from the REFERENCE'S OWN SOURCE.

oracle/_ref/ref_dump is /root/reference/src/QPSolver.cpp compiled unmodified (recipe: `make -C oracle ref`; Eigen and
qpOASES resolved to the stand-ins under oracle/ref_shim/) plus a small driver (oracle/ref_dump.cpp).  Every array
stored under a "ref_" key below was computed by that binary: QPSolver::discretizeSystem (QPSolver.cpp:21-29),
buildQPParams (:31-81) and updateState (:108-111).  Inputs are stored beside them so the tests feed the oracle and the
CUDA path the same bytes.  The box that runs the GPU tier has no /root/reference: it only reads the committed .npz.

What this pins: a6 (Ad, Bd), a7 (A_aug x0 through b_eq, B_aug through A_eq), a8 (H, f), a9 generic rows (lb, ub, A_ineq,
lbA, ubA), a11 (x+).  What it does not: qpOASES's arithmetic (absent; the demo loop below closes through the oracle's
active-set solver behind the qpOASES-shaped shim), the TRON1 "intended physics" model matrices themselves (they are
inputs here; include/mpcQP.h does not compile), the per-step (ltv=1) extension.
Run (in the container that has /root/reference):  python tests/golden/make_ref_golden.py
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")


def build_ref():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libmpc_oracle.so", "ref"])
    return DUMP


def cases():
    """name -> dict of inputs.  Model matrices for the TRON1 cases come from the oracle (they are INPUTS to the
    reference code here, not results)."""
    import oracle_lib as O
    from mpc_limx_control_b200 import synth
    out = {}
    # the reference demo system (src/qpSolver_test.cpp:6-24), three (x0, reference) pairs
    Ac = np.array([[0, 1, 0, 0], [0, -0.1, 0, 0], [0, 0, 0, 1], [0, 0, 0, -0.1]], dtype=float)
    Bc = np.array([[0, 0], [5, 0], [0, 0], [0, 5]], dtype=float)
    Q = np.diag([50.0, 5, 50, 5]); R = 0.1 * np.eye(2); P = 20 * Q
    xm = np.array([-5.0, -3, -5, -3])

    def circle(k, Ts, N):
        t = (k + np.arange(N + 1)) * Ts
        th = 0.5 * t
        return np.stack([2 * np.cos(th), -2 * 0.5 * np.sin(th), 2 * np.sin(th), 2 * 0.5 * np.cos(th)])
    for nm, k, x0, ub in [("demo0", 0, [2.0, 0, 0, 0], 8.0), ("demo40", 40, [2.0, 0.5, -1.0, 0.2], 2.0),
                          ("demo300", 300, [-1.2, 0.4, 1.7, -0.9], 8.0)]:
        out[nm] = dict(NX=4, NU=2, N=15, Ts=0.01, u_min=-ub, u_max=ub, Ac=Ac, Bc=Bc, Q=Q, R=R, P=P, x_min=xm, x_max=-xm,
                       xi0=np.array(x0), xi_ref=circle(k, 0.01, 15), u=np.array([0.3, -0.7]))
    # random systems, header-typed sizes (strict fixed-size mode) and others (relaxed mode), incl. expm branches
    rng = np.random.default_rng(20261018)
    for nm, NX, NU, N, Ts, sc in [("rnd4a", 4, 2, 15, 0.01, 1.0), ("rnd4b", 4, 2, 15, 0.2, 6.0), ("rnd4c", 4, 2, 7, 1.0, 9.0),
                                  ("rnd6", 6, 3, 8, 0.05, 3.0), ("rnd9", 9, 1, 12, 0.3, 2.0)]:
        A = rng.standard_normal((NX, NX)) * sc - sc * np.eye(NX); B = rng.standard_normal((NX, NU))
        M = rng.standard_normal((NX, NX)); Qr = M @ M.T + np.eye(NX)
        M = rng.standard_normal((NU, NU)); Rr = M @ M.T + 0.1 * np.eye(NU)
        out[nm] = dict(NX=NX, NU=NU, N=N, Ts=Ts, u_min=-3.0, u_max=4.0, Ac=A, Bc=B, Q=Qr, R=Rr, P=5 * Qr,
                       x_min=-rng.uniform(1, 5, NX), x_max=rng.uniform(1, 5, NX), xi0=rng.standard_normal(NX),
                       xi_ref=rng.standard_normal((NX, N + 1)), u=rng.standard_normal(NU))
    # TRON1: the intended 13 x 6 model at synthetic instances (one model at x0 = the reference's LTI structure)
    big = np.full(13, 1.0e3)
    for nm, N, Ts, seed, idx in [("tron10a", 10, 0.005, 1001, 0), ("tron10b", 10, 0.005, 1001, 7), ("tron10stiff", 10, 0.05, 4, 1),
                                 ("tron20", 20, 0.005, 1002, 3), ("tron50", 50, 0.005, 1003, 5)]:
        b = synth.tron1_batch(seed, idx + 1, N, Ts)
        p = O.tron1_defaults(Ts=Ts)
        x0 = b["x0"][idx].copy(); feet = b["feet"][idx].copy()
        Ac13, Bc13 = O.tron1_model(p, float(x0[2]), x0[3:6], feet.reshape(-1))
        Q13 = np.diag(np.array(p.q[:])); R6 = p.r * np.eye(6)
        out[nm] = dict(NX=13, NU=6, N=N, Ts=Ts, u_min=-200.0, u_max=200.0, Ac=Ac13, Bc=Bc13, Q=Q13, R=R6, P=p.p_scale * Q13,
                       x_min=-big, x_max=big, xi0=x0, xi_ref=np.ascontiguousarray(b["x_ref"][idx].T), u=rng.uniform(-20, 60, 6),
                       feet=feet)
    # TRON1 reference-literal 13 x 3 model (include/mpcQP.h:139-181), mpcQP's own Ts and horizon (:37-38)
    p = O.tron1_defaults()
    AcL, BcL = O.tron1_model_literal(p, np.array([0.1, 0.2, 0.8]), np.array([0.05, 0.1, 0.0]))
    x0 = np.array([0.01, 0.02, 0.3, 0.1, 0.2, 0.8, 0.0, 0.1, 0.0, 0.3, 0.0, 0.0, -9.8])
    Q13 = np.diag(np.array(p.q[:]))
    out["literal"] = dict(NX=13, NU=3, N=20, Ts=0.001, u_min=-8.0, u_max=8.0, Ac=AcL, Bc=BcL, Q=Q13, R=0.1 * np.eye(3), P=20 * Q13,
                          x_min=-big, x_max=big, xi0=x0, xi_ref=O.tron1_reference(x0, 20, 0.001), u=np.array([1.0, -2.0, 3.0]))
    return out


def F(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64)).tobytes(order="F")


def run_cases(cs):
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<i", len(cs)))
            for c in cs.values():
                f.write(struct.pack("<iii", c["NX"], c["NU"], c["N"]))
                f.write(struct.pack("<ddd", c["Ts"], c["u_min"], c["u_max"]))
                for k in ("Ac", "Bc", "Q", "R", "P", "x_min", "x_max", "xi0", "xi_ref", "u"):
                    f.write(F(c[k]))
        subprocess.check_call([DUMP, "cases", fin, fout], stderr=subprocess.DEVNULL)
        a = np.fromfile(fout)
    res = {}
    o = 0
    for nm, c in cs.items():
        NX, NU, N = c["NX"], c["NU"], c["N"]
        n = NU * N
        r = {}
        for k, shp in [("Ad", (NX, NX)), ("Bd", (NX, NU)), ("H", (n, n)), ("f", (n,)), ("A_eq", (NX * N, n)), ("b_eq", (NX * N,)),
                       ("lb", (n,)), ("ub", (n,)), ("A_ineq", (2 * NX * N, n)), ("lbA", (2 * NX * N,)), ("ubA", (2 * NX * N,)),
                       ("x_next", (NX,))]:
            cnt = int(np.prod(shp))
            r[k] = a[o:o + cnt].reshape(shp, order="F").copy(); o += cnt
        res[nm] = r
    assert o == a.size, (o, a.size)
    return res


def run_demo():
    with tempfile.TemporaryDirectory() as td:
        fout = os.path.join(td, "demo.bin")
        subprocess.check_call([DUMP, "demo", fout], stderr=subprocess.DEVNULL)
        a = np.fromfile(fout)
    o = 0

    def take(*shp):
        nonlocal o
        cnt = int(np.prod(shp)); v = a[o:o + cnt].reshape(shp, order="F").copy(); o += cnt
        return v
    d = dict(Ad=take(4, 4), Bd=take(4, 2))
    xs = [take(4)]; us = []; st = []
    for k in range(500):
        us.append(take(2)); xs.append(take(4)); st.append(take(1)[0])
        if k == 0:
            d["as_written_status"] = take(1)[0]
            d.update(H0=take(30, 30), f0=take(30), A_ineq0=take(120, 30), lbA0=take(120), ubA0=take(120), U0=take(2, 15))
    assert o == a.size
    d.update(xs=np.array(xs), us=np.array(us), status=np.array(st))
    return d


def generate():
    build_ref()
    cs = cases()
    res = run_cases(cs)
    out = {"case_names": np.array(list(cs.keys()))}
    for nm, c in cs.items():
        for k, v in c.items():
            out[f"in_{nm}_{k}"] = np.asarray(v)
        for k, v in res[nm].items():
            out[f"ref_{nm}_{k}"] = v
    for k, v in run_demo().items():
        out[f"ref_demo_{k}"] = np.asarray(v)
    return out


if __name__ == "__main__":
    out = generate()
    path = os.path.join(HERE, "ref_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e3:.0f} kB,", len(out), "arrays")
