"""GPU tier (-m gpu): the CUDA path through the C ABI against the oracle on the same seeded inputs,
against the committed golden fixtures, and -- at BASELINE.json's full batch sizes -- through
size-independent properties.  Tolerances (BASELINE.json north_star): H, f, prediction matrices
1e-9 relative in FP64; forces 1e-4 relative with KKT (natural) residual <= 1e-6; contact/mode
indices bit-exact."""
import os

import numpy as np
import pytest

import oracle_lib as O
from mpc_limx_control_b200 import synth

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max())


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    return torch


def make_engine(N, B, **kw):
    from mpc_limx_control_b200.engine import Engine
    return Engine(horizon=N, max_batch=B, device=0, **kw)


def to_dev(torch, d, keys=("x0", "x_ref", "feet", "iter")):
    return {k: torch.from_numpy(np.ascontiguousarray(d[k])).cuda() for k in keys}


def test_library_loaded_and_device(torch_cuda):
    from mpc_limx_control_b200 import _capi
    assert _capi.lib().mpc_b200_device_count() >= 1


def test_golden_cases(torch_cuda, golden):
    torch = torch_cuda
    for key in golden["tron1_cases"]:
        N = int(golden[f"{key}_N"]); Ts = float(golden[f"{key}_Ts"]); ltv = int(golden[f"{key}_ltv"])
        eng = make_engine(N, 4, Ts=Ts, ltv=ltv)
        x0 = torch.from_numpy(golden[f"{key}_x0"][None].copy()).cuda()
        xr = torch.from_numpy(golden[f"{key}_xref"].T[None].copy()).cuda()
        feet = torch.from_numpy(golden[f"{key}_feet"][None].copy()).cuda()
        contact = torch.from_numpy(golden[f"{key}_contact"][None].copy()).cuda()
        c = eng.condense(x0, xr, feet)
        for k in ("H", "f", "A_aug", "B_aug"):
            assert rel(c[k][0], golden[f"{key}_{k}"]) < 1e-9, (key, k)
        F, st, it = eng.solve(x0, xr, feet, contact=contact)
        torch.cuda.synchronize()
        assert int(st[0]) == 0, key
        Fn = F.cpu().numpy()[0]
        assert np.abs(Fn.reshape(-1) - golden[f"{key}_U"]).max() / max(1.0, np.abs(Fn).max()) < 1e-5, key
        eng.close()


@pytest.mark.parametrize("N,Ts,ltv,standing,scale,mu,B", [
    (10, 0.005, 1, False, 1, 0.5, 67),     # config-2 distribution, odd batch (tail CTA, non-bulk path)
    (10, 0.005, 0, False, 1, 0.5, 32),     # LTI (reference QPSolver structure)
    (10, 0.005, 1, True, 1, 0.5, 32),      # standing, n = 60 free variables
    (10, 0.05, 1, False, 3, 0.5, 32),      # ill-conditioned
    (10, 0.001, 1, False, 1, 0.5, 32),     # mpcQP.h Ts, fz >= 0 active
    (10, 0.02, 1, True, 8, 0.2, 48),       # friction pyramid heavily active
    (20, 0.005, 1, False, 1, 0.5, 21),     # config-3 horizon
    (20, 0.05, 1, True, 6, 0.3, 16),       # N=20 stress
    (50, 0.005, 1, False, 1, 0.5, 5),      # config-4 horizon, single stance (n = 150, factor in shared memory)
    (50, 0.01, 1, True, 3, 0.4, 3),        # config-4 horizon, double support (n = 300, factor in global memory)
])
def test_batch_vs_oracle(torch_cuda, N, Ts, ltv, standing, scale, mu, B):
    torch = torch_cuda
    d = synth.tron1_batch(77, B, N, Ts, standing=standing)
    d["x0"][:, [0, 1, 6, 7, 8, 9, 10, 11]] *= scale
    eng = make_engine(N, B, Ts=Ts, ltv=ltv, mu=mu)
    t = to_dev(torch, d)
    contact = eng.contact_schedule(t["iter"])
    c_ref = np.stack([O.contact_schedule(int(i), N) for i in d["iter"]])
    assert np.array_equal(contact.cpu().numpy(), c_ref)                       # bit-exact
    po = O.tron1_defaults(Ts=Ts, ltv=ltv, mu=mu)
    dump = eng.condense(t["x0"], t["x_ref"], t["feet"])
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], contact=contact)
    F2, st2, it2 = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])    # in-kernel gait schedule
    torch.cuda.synchronize()
    F = F.cpu().numpy(); st = st.cpu().numpy()
    assert np.array_equal(F, F2.cpu().numpy()) and np.array_equal(st, st2.cpu().numpy())
    assert (st == 0).all(), st
    Fo, so, _ = O.tron1_solve_batch(po, N, d["x0"], d["x_ref"], d["feet"], c_ref, nthreads=8)
    assert (so == 0).all()
    for b in range(B):
        c = O.tron1_condense(po, N, d["x0"][b], d["x_ref"][b], d["feet"][b], want_pred=(b < 4))
        assert rel(dump["H"][b], c["H"]) < 1e-9 and rel(dump["f"][b], c["f"]) < 1e-9
        if b < 4:
            assert rel(dump["A_aug"][b], c["A_aug"]) < 1e-9 and rel(dump["B_aug"][b], c["B_aug"]) < 1e-9
        assert np.abs(F[b] - Fo[b]).max() / max(1.0, np.abs(Fo[b]).max()) < 1e-4
        assert O.tron1_natural_residual(po, N, c["H"], c["f"], c_ref[b], F[b]) < 1e-6
        assert np.all(F[b].reshape(N, 2, 3)[c_ref[b] == 0] == 0.0)
    eng.close()


def test_host_entry_point_matches_device(torch_cuda):
    torch = torch_cuda
    N, B, Ts = 10, 130, 0.005
    d = synth.tron1_batch(5, B, N, Ts)
    eng = make_engine(N, B, Ts=Ts)
    t = to_dev(torch, d)
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize()
    pin = {k: torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory() for k in ("x0", "x_ref", "feet", "iter")}
    Fh = torch.empty((B, N, 6), dtype=torch.float64).pin_memory()
    sh = torch.empty(B, dtype=torch.int32).pin_memory(); ih = torch.empty(B, dtype=torch.int32).pin_memory()
    eng.solve_host(pin["x0"], pin["x_ref"], pin["feet"], it=pin["iter"], forces=Fh, status=sh, iters=ih)
    assert np.array_equal(Fh.numpy(), F.cpu().numpy()) and np.array_equal(sh.numpy(), st.cpu().numpy())
    # pageable numpy buffers and an explicit contact schedule
    contact = np.stack([O.contact_schedule(int(i), N) for i in d["iter"]])
    F3, s3, i3 = eng.solve_host(d["x0"], d["x_ref"], d["feet"], contact=contact)
    assert np.array_equal(F3, F.cpu().numpy())
    eng.close()


def test_per_step_feet(torch_cuda):
    torch = torch_cuda
    N, B, Ts = 10, 24, 0.005
    d = synth.tron1_batch(8, B, N, Ts, per_step_feet=True)
    rng = np.random.default_rng(0)
    d["feet"] = d["feet"] + rng.uniform(-0.02, 0.02, d["feet"].shape) * np.array([1, 1, 0.0])
    eng = make_engine(N, B, Ts=Ts, per_step_feet=1)
    t = to_dev(torch, d)
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize()
    po = O.tron1_defaults(Ts=Ts, per_step_feet=1)
    c_ref = np.stack([O.contact_schedule(int(i), N) for i in d["iter"]])
    Fo, so, _ = O.tron1_solve_batch(po, N, d["x0"], d["x_ref"], d["feet"], c_ref, nthreads=8)
    assert (st.cpu().numpy() == 0).all() and (so == 0).all()
    assert np.abs(F.cpu().numpy() - Fo).max() / max(1.0, np.abs(Fo).max()) < 1e-4
    eng.close()


def test_edge_batches(torch_cuda):
    """B = 1, 2, 3, 5 (ragged CTAs) and an all-swing instance."""
    torch = torch_cuda
    N, Ts = 10, 0.005
    eng = make_engine(N, 8, Ts=Ts)
    po = O.tron1_defaults(Ts=Ts)
    for B in (1, 2, 3, 5):
        d = synth.tron1_batch(100 + B, B, N, Ts)
        t = to_dev(torch, d)
        c_ref = np.stack([O.contact_schedule(int(i), N) for i in d["iter"]])
        F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
        torch.cuda.synchronize()
        Fo, so, _ = O.tron1_solve_batch(po, N, d["x0"], d["x_ref"], d["feet"], c_ref)
        assert (st.cpu().numpy() == 0).all()
        assert np.abs(F.cpu().numpy() - Fo).max() / max(1.0, np.abs(Fo).max()) < 1e-4
    d = synth.tron1_batch(3, 2, N, Ts)
    t = to_dev(torch, d)
    zero = torch.zeros((2, N, 2), dtype=torch.uint8, device="cuda")
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], contact=zero)
    torch.cuda.synchronize()
    assert (st.cpu().numpy() == 0).all() and float(F.abs().max()) == 0.0
    eng.close()


def test_iteration_cap_status(torch_cuda):
    torch = torch_cuda
    N, Ts, B = 10, 0.02, 16
    d = synth.tron1_batch(9, B, N, Ts, standing=True)
    d["x0"][:, [0, 1, 6, 7, 8, 9, 10, 11]] *= 5
    eng = make_engine(N, B, Ts=Ts, max_newton=1, max_admm=2)
    t = to_dev(torch, d)
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize()
    st = st.cpu().numpy(); F = F.cpu().numpy()
    assert (st == 1).any() and np.isfinite(F).all()
    F3 = F.reshape(B, N, 2, 3)
    assert (np.abs(F3[..., 0]) <= 0.5 * F3[..., 2] + 1e-9).all()
    eng.close()


def test_capacity_and_argument_errors(torch_cuda):
    torch = torch_cuda
    from mpc_limx_control_b200 import _capi
    eng = make_engine(10, 4)
    d = synth.tron1_batch(1, 8, 10, 0.005)
    with pytest.raises(_capi.MpcB200Error) as ei:
        eng.solve_host(d["x0"], d["x_ref"], d["feet"], it=d["iter"])
    assert ei.value.code == _capi.ECAPACITY
    t = to_dev(torch, d)
    with pytest.raises(_capi.MpcB200Error):
        eng.solve(t["x0"], t["x_ref"], t["feet"])            # neither contact nor iter
    eng.close()


def test_full_size_properties_config2(torch_cuda):
    """BASELINE config 2 (B=4096, N=10, seed 1001): size-independent properties at full size and
    oracle parity on a strided sample."""
    torch = torch_cuda
    N, B, Ts = 10, 4096, 0.005
    d = synth.tron1_batch(1001, B, N, Ts)
    eng = make_engine(N, B, Ts=Ts)
    t = to_dev(torch, d)
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    Fb, stb, _ = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize()
    F = F.cpu().numpy(); st = st.cpu().numpy()
    assert np.array_equal(F, Fb.cpu().numpy())                    # deterministic / idempotent
    assert (st == 0).all()
    c_ref = np.stack([O.contact_schedule(int(i), N) for i in d["iter"]])
    F4 = F.reshape(B, N, 2, 3)
    assert np.all(F4[c_ref == 0] == 0.0)                          # swing feet: exactly zero force
    assert (np.abs(F4[..., 0]) <= 0.5 * F4[..., 2] + 1e-9).all()  # friction pyramid
    assert (np.abs(F4[..., 1]) <= 0.5 * F4[..., 2] + 1e-9).all()
    assert (F4[..., 2] >= -1e-12).all() and (F4[..., 2] <= 2 * 9.585 * 9.8 + 1e-9).all()
    # permutation equivariance: solving a shuffled batch gives the shuffled forces
    perm = np.random.default_rng(0).permutation(B)
    tp = {k: t[k][torch.from_numpy(perm).cuda()].contiguous() for k in t}
    Fp, _, _ = eng.solve(tp["x0"], tp["x_ref"], tp["feet"], it=tp["iter"])
    torch.cuda.synchronize()
    assert np.array_equal(Fp.cpu().numpy(), F[perm])
    po = O.tron1_defaults(Ts=Ts)
    idx = np.arange(0, B, 64)
    Fo, so, _ = O.tron1_solve_batch(po, N, d["x0"][idx], d["x_ref"][idx], d["feet"][idx], c_ref[idx], nthreads=8)
    assert (so == 0).all()
    assert np.abs(F[idx] - Fo).max() / max(1.0, np.abs(Fo).max()) < 1e-4
    eng.close()


def _check_force_properties(F, c_ref, mu=0.5, fmax=2 * 9.585 * 9.8):
    B, N = F.shape[0], F.shape[1]
    F4 = F.reshape(B, N, 2, 3)
    assert np.all(F4[c_ref == 0] == 0.0)                              # swing feet: exactly zero force
    assert (np.abs(F4[..., 0]) <= mu * F4[..., 2] + 1e-9).all()       # friction pyramid
    assert (np.abs(F4[..., 1]) <= mu * F4[..., 2] + 1e-9).all()
    assert (F4[..., 2] >= -1e-12).all() and (F4[..., 2] <= fmax + 1e-9).all()


@pytest.mark.parametrize("name,N,B,seed,stride", [("config3", 20, 65536, 1002, 2048), ("config4", 50, 8192, 1003, 2048)])
def test_full_size_properties_configs_3_4(torch_cuda, name, N, B, seed, stride):
    """BASELINE configs[2] (B=65536, N=20) and configs[3] (B=8192, N=50) at full size on one GPU: every solve
    certified, feasibility, exact zeros on swing feet, determinism, and oracle parity on a strided sample."""
    torch = torch_cuda
    Ts = 0.005
    d = synth.tron1_batch(seed, B, N, Ts)
    eng = make_engine(N, B, Ts=Ts)
    t = to_dev(torch, d)
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    Fb, _, _ = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize()
    assert torch.equal(F, Fb) and int((st != 0).sum()) == 0
    idx = np.arange(0, B, stride)
    Fs = F[torch.from_numpy(idx).cuda()].cpu().numpy()
    contact = eng.contact_schedule(t["iter"]).cpu().numpy()
    c_ref = np.stack([O.contact_schedule(int(i), N) for i in d["iter"][idx]])
    assert np.array_equal(contact[idx], c_ref)                        # mode indices bit-exact
    _check_force_properties(F.cpu().numpy(), contact)
    po = O.tron1_defaults(Ts=Ts)
    Fo, so, _ = O.tron1_solve_batch(po, N, d["x0"][idx], d["x_ref"][idx], d["feet"][idx], c_ref, nthreads=8)
    assert (so == 0).all()
    assert np.abs(Fs - Fo).max() / max(1.0, np.abs(Fo).max()) < 1e-4
    eng.close()


def test_full_size_rollout_config5(torch_cuda):
    """BASELINE configs[4]: 16384 robots x 1000 closed-loop control steps (one GPU holds the whole job here):
    every step certified, states finite, forces feasible, and the first 25 steps of a strided sample equal the
    oracle loop."""
    torch = torch_cuda
    N, Ts, steps, B = 10, 0.005, 1000, 16384
    d = synth.tron1_batch(1004, B, N, Ts)
    eng = make_engine(N, B, Ts=Ts)
    x = torch.from_numpy(d["x0"].copy()).cuda()
    args = (torch.from_numpy(d["omega_yaw"]).cuda(), torch.from_numpy(d["velocity_x"]).cuda(), torch.from_numpy(d["iter"]).cuda())
    _, bad, its = eng.rollout(x, *args, steps)
    torch.cuda.synchronize()
    assert int(bad.sum()) == 0 and int(its.min()) >= steps and bool(torch.isfinite(x).all())
    idx = np.arange(0, B, 4096)
    xs = torch.from_numpy(d["x0"][idx].copy()).cuda()
    traj, bad2, _ = eng.rollout(xs, args[0][idx].contiguous(), args[1][idx].contiguous(), args[2][idx].contiguous(), 25, want_traj=True)
    torch.cuda.synchronize()
    U = traj.cpu().numpy(); X = xs.cpu().numpy()
    assert (U[..., 2] >= -1e-12).all() and (U[..., 5] >= -1e-12).all() and (np.abs(U[..., 0]) <= 0.5 * U[..., 2] + 1e-9).all()
    po = O.tron1_defaults(Ts=Ts)
    offl = list(eng.params.foot_offset_left); offr = list(eng.params.foot_offset_right)
    for j, b in enumerate(idx):
        xo, Uo, bo = O.tron1_rollout(po, N, 25, d["x0"][b], d["omega_yaw"][b], d["velocity_x"][b], int(d["iter"][b]), offl, offr)
        assert bo == 0 and np.abs(U[j] - Uo).max() / max(1.0, np.abs(Uo).max()) < 1e-4 and np.abs(X[j] - xo).max() < 1e-6
    eng.close()


@pytest.mark.parametrize("seed,Ts,scale,mu,standing_every", [(9001, 0.005, 1.0, 0.5, 0), (9002, 0.02, 4.0, 0.35, 7), (9003, 0.05, 2.0, 0.25, 0)])
def test_fuzz_forces_vs_oracle(torch_cuda, seed, Ts, scale, mu, standing_every):
    """16,384 random instances per setting, EVERY one compared with the oracle's active-set solution: nominal, a
    stressed setting where most instances need several active-face iterations (and some the ADMM fallback), and the
    ill-conditioned Ts = 0.05 (cond(H) ~ 3e4).  Tolerances of BASELINE.json: forces 1e-4 relative, natural residual 1e-6."""
    torch = torch_cuda
    N, B = 10, 16384
    d = synth.tron1_batch(seed, B, N, Ts)
    d["x0"][:, [0, 1, 6, 7, 8, 9, 10, 11]] *= scale
    if standing_every:
        d["iter"][::standing_every] = -1
    eng = make_engine(N, B, Ts=Ts, mu=mu)
    t = to_dev(torch, d)
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize()
    F = F.cpu().numpy(); st = st.cpu().numpy(); it = it.cpu().numpy()
    assert (st == 0).all(), np.bincount(st)
    c_ref = np.stack([O.contact_schedule(int(i), N) for i in d["iter"]])
    po = O.tron1_defaults(Ts=Ts, mu=mu)
    Fo, so, _ = O.tron1_solve_batch(po, N, d["x0"], d["x_ref"], d["feet"], c_ref, nthreads=os.cpu_count() or 8)
    assert (so == 0).all()
    err = np.abs(F - Fo).reshape(B, -1).max(1) / np.maximum(1.0, np.abs(Fo).reshape(B, -1).max(1))
    assert err.max() < 1e-4, (err.max(), int(err.argmax()), int(it[err.argmax()]))
    _check_force_properties(F, c_ref, mu=mu)
    if scale > 1.0:
        assert it.max() >= 3 and (it > 1).mean() > 0.05        # the setting really exercises the iterations
    eng.close()


def test_fp64_peak_measurement(torch_cuda):
    from mpc_limx_control_b200.engine import measure_fp64_peak
    tf = measure_fp64_peak(0)
    assert 5.0 < tf < 80.0, tf


def test_reference_generator(torch_cuda):
    """mpcQP's reference generator on the device == synth's numpy restatement of include/mpcQP.h:74-97"""
    torch = torch_cuda
    for N in (10, 20):
        d = synth.tron1_batch(12, 257, N, 0.005)
        eng = make_engine(N, 257)
        xr = eng.reference(torch.from_numpy(d["x0"]).cuda(), torch.from_numpy(d["omega_yaw"]).cuda(),
                           torch.from_numpy(d["velocity_x"]).cuda())
        torch.cuda.synchronize()
        assert np.abs(xr.cpu().numpy() - d["x_ref"]).max() < 1e-14    # FMA contraction: last-bit differences only
        eng.close()


def test_rollout_vs_oracle(torch_cuda):
    """closed-loop rollout (BASELINE configs[4] shape): device loop vs the oracle loop, incl. a standing
    instance (large capacity class) and one whose gait switches support foot mid-rollout"""
    torch = torch_cuda
    N, Ts, steps, B = 10, 0.005, 40, 9
    d = synth.tron1_batch(1004, B, N, Ts)
    it0 = d["iter"].copy(); it0[1] = 480; it0[2] = -1; it0[5] = 995
    eng = make_engine(N, B, Ts=Ts)
    x = torch.from_numpy(d["x0"].copy()).cuda()
    traj, bad, its = eng.rollout(x, torch.from_numpy(d["omega_yaw"]).cuda(), torch.from_numpy(d["velocity_x"]).cuda(),
                                 torch.from_numpy(it0).cuda(), steps, want_traj=True)
    torch.cuda.synchronize()
    assert int(bad.sum()) == 0 and int(its.min()) >= steps
    po = O.tron1_defaults(Ts=Ts)
    offl = list(eng.params.foot_offset_left); offr = list(eng.params.foot_offset_right)
    X = x.cpu().numpy(); U = traj.cpu().numpy()
    for b in range(B):
        xo, Uo, bo = O.tron1_rollout(po, N, steps, d["x0"][b], d["omega_yaw"][b], d["velocity_x"][b], int(it0[b]), offl, offr)
        assert bo == 0
        assert np.abs(U[b] - Uo).max() / max(1.0, np.abs(Uo).max()) < 1e-4
        assert np.abs(X[b] - xo).max() < 1e-6
    eng.close()


def test_rollout_properties_larger(torch_cuda):
    torch = torch_cuda
    N, Ts, steps, B = 10, 0.005, 200, 1024
    d = synth.tron1_batch(1004, B, N, Ts)
    eng = make_engine(N, B, Ts=Ts)
    x = torch.from_numpy(d["x0"].copy()).cuda(); x2 = x.clone()
    args = (torch.from_numpy(d["omega_yaw"]).cuda(), torch.from_numpy(d["velocity_x"]).cuda(), torch.from_numpy(d["iter"]).cuda())
    traj, bad, its = eng.rollout(x, *args, steps, want_traj=True)
    # two half rollouts == one full rollout (state fully carried; the clock advances by steps*mpc_step)
    _, bad2a, _ = eng.rollout(x2, *args, steps // 2)
    it_half = (torch.from_numpy(d["iter"]) + (steps // 2) * 5).to(torch.int32).cuda()
    _, bad2b, _ = eng.rollout(x2, args[0], args[1], it_half, steps // 2)
    torch.cuda.synchronize()
    assert int(bad.sum()) == 0 and int(bad2a.sum()) == 0 and int(bad2b.sum()) == 0
    X = x.cpu().numpy(); U = traj.cpu().numpy()
    assert np.isfinite(X).all() and np.isfinite(U).all()
    assert np.abs(X - x2.cpu().numpy()).max() < 1e-9
    assert (U[..., 2] >= -1e-12).all() and (U[..., 5] >= -1e-12).all()
    assert (np.abs(U[..., 0]) <= 0.5 * U[..., 2] + 1e-9).all()
    eng.close()


def test_host_paths_packed_and_pipelined(torch_cuda):
    """solve_host small-batch packed path (B<=64) and chunk-pipelined path (B>=1024, ragged tail),
    and the controller-shaped entry (command in, first-step force out), all bit-identical to the
    device entry point on the same inputs."""
    torch = torch_cuda
    from mpc_limx_control_b200.engine import control_host
    N, Ts = 10, 0.005
    for B in (1, 7, 64, 65, 2050, 6001):
        d = synth.tron1_batch(21, B, N, Ts)
        if B == 7:
            d["iter"][3] = -1          # one standing instance: exercises the overflow class in the packed path
        eng = make_engine(N, max(B, 8), Ts=Ts)
        t = to_dev(torch, d)
        F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
        torch.cuda.synchronize()
        F = F.cpu().numpy(); st = st.cpu().numpy()
        assert (st == 0).all()
        Fh, sh, ih = eng.solve_host(d["x0"], d["x_ref"], d["feet"], it=d["iter"])
        assert np.array_equal(Fh, F) and np.array_equal(sh, st)
        contact = np.stack([O.contact_schedule(int(i), N) for i in d["iter"]])
        Fh2, sh2, _ = eng.solve_host(d["x0"], d["x_ref"], d["feet"], contact=contact)
        assert np.array_equal(Fh2, F)
        u0, s0, i0 = control_host(eng, d["x0"], d["omega_yaw"], d["velocity_x"], d["feet"], it=d["iter"])
        assert (s0 == 0).all()
        # the device reference generator may differ from numpy's x_ref in the last bit (FMA) -> tiny tolerance
        assert np.abs(u0 - F[:, 0, :]).max() / max(1.0, np.abs(F).max()) < 1e-9
        eng.close()


def test_host_zero_copy_and_staged_paths_agree(torch_cuda):
    """Pinned buffers take the zero-copy path (kernel reads/writes host memory directly), pageable ones the
    staged path; HOST_ZEROCOPY refuses pageable buffers; both paths are bit-identical to the device entry point.
    Covers a standing instance (overflow class reading host memory through the list-driven kernel) and an
    8-byte-aligned (not 16) slice so that the plain-load staging fallback runs against host memory too."""
    torch = torch_cuda
    from mpc_limx_control_b200.engine import Engine, control_host
    from mpc_limx_control_b200 import _capi
    N, Ts = 10, 0.005
    for B in (1, 5, 300, 4100):
        d = synth.tron1_batch(33, B + 1, N, Ts)
        d["iter"][min(2, B - 1)] = -1
        eng = make_engine(N, B + 1, Ts=Ts)
        t = to_dev(torch, {k: v[:B] for k, v in d.items()})
        F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
        torch.cuda.synchronize()
        F = F.cpu().numpy(); st = st.cpu().numpy(); it = it.cpu().numpy()
        pin = {k: torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory() for k in ("x0", "x_ref", "feet", "iter", "omega_yaw", "velocity_x")}
        Fh = torch.zeros((B, N, 6), dtype=torch.float64).pin_memory()
        sh = torch.full((B,), -7, dtype=torch.int32).pin_memory(); ih = torch.zeros(B, dtype=torch.int32).pin_memory()
        for off in (0, 1):            # off = 1: instance slices start 8-byte aligned only
            Fh.zero_(); sh.fill_(-7)
            sl = {k: v[off:off + B] for k, v in pin.items()}
            eng.set_host_mode(Engine.HOST_ZEROCOPY)
            eng.solve_host(sl["x0"], sl["x_ref"], sl["feet"], it=sl["iter"], forces=Fh, status=sh, iters=ih)
            assert eng.last_host_path() == 1
            t2 = to_dev(torch, {k: d[k][off:off + B] for k in ("x0", "x_ref", "feet", "iter")})
            F2, st2, it2 = eng.solve(t2["x0"], t2["x_ref"], t2["feet"], it=t2["iter"])
            torch.cuda.synchronize()
            assert np.array_equal(Fh.numpy(), F2.cpu().numpy()) and np.array_equal(sh.numpy(), st2.cpu().numpy())
            assert np.array_equal(ih.numpy(), it2.cpu().numpy())
        # pageable buffers: AUTO falls back to staged copies, ZEROCOPY refuses
        eng.set_host_mode(Engine.HOST_AUTO)
        Fp, sp, ip = eng.solve_host(d["x0"][:B], d["x_ref"][:B], d["feet"][:B], it=d["iter"][:B])
        assert eng.last_host_path() == 0 and np.array_equal(Fp, F) and np.array_equal(sp, st) and np.array_equal(ip, it)
        eng.set_host_mode(Engine.HOST_ZEROCOPY)
        with pytest.raises(_capi.MpcB200Error):
            eng.solve_host(d["x0"][:B], d["x_ref"][:B], d["feet"][:B], it=d["iter"][:B])
        # pinned + forced staging == zero-copy, and the controller-shaped entry agrees across paths
        eng.set_host_mode(Engine.HOST_STAGED)
        Fs = torch.zeros((B, N, 6), dtype=torch.float64).pin_memory()
        eng.solve_host(pin["x0"][:B], pin["x_ref"][:B], pin["feet"][:B], it=pin["iter"][:B], forces=Fs, status=sh, iters=ih)
        assert eng.last_host_path() == 0 and np.array_equal(Fs.numpy(), F)
        u_s = torch.zeros((B, 6), dtype=torch.float64).pin_memory(); u_z = torch.zeros((B, 6), dtype=torch.float64).pin_memory()
        control_host(eng, pin["x0"][:B], pin["omega_yaw"][:B], pin["velocity_x"][:B], pin["feet"][:B], it=pin["iter"][:B], u0=u_s, status=sh, iters=ih)
        eng.set_host_mode(Engine.HOST_ZEROCOPY)
        control_host(eng, pin["x0"][:B], pin["omega_yaw"][:B], pin["velocity_x"][:B], pin["feet"][:B], it=pin["iter"][:B], u0=u_z, status=sh, iters=ih)
        assert eng.last_host_path() == 1 and np.array_equal(u_s.numpy(), u_z.numpy())
        eng.close()


def test_host_paths_all_double_support_batches(torch_cuda):
    """when the schedule in host memory shows that EVERY instance is double support (standing robots), the host entry
    points launch the large-class kernel alone over the identity list: same bits as the device entry point, for the
    gait-clock and the explicit-contact forms, pinned (zero-copy) and pageable (staged), horizons 10 and 20"""
    torch = torch_cuda
    for N in (10, 20):
        for B in (1, 6, 300):
            d = synth.tron1_batch(91, B, N, 0.005, standing=True)
            eng = make_engine(N, max(B, 8), Ts=0.005)
            t = to_dev(torch, d)
            F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
            torch.cuda.synchronize()
            F = F.cpu().numpy(); st = st.cpu().numpy()
            assert (st == 0).all() and (d["iter"] < 0).all()
            l0 = eng.launch_count()
            Fh, sh, _ = eng.solve_host(d["x0"], d["x_ref"], d["feet"], it=d["iter"])                  # pageable
            assert np.array_equal(Fh, F) and np.array_equal(sh, st)
            contact = np.ones((B, N, 2), np.uint8)
            pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in dict(d, contact=contact).items() if k in ("x0", "x_ref", "feet", "contact")}
            Fp = torch.zeros((B, N, 6), dtype=torch.float64).pin_memory(); sp = torch.zeros(B, dtype=torch.int32).pin_memory()
            ip = torch.zeros(B, dtype=torch.int32).pin_memory()
            l1 = eng.launch_count()
            eng.solve_host(pin["x0"], pin["x_ref"], pin["feet"], contact=pin["contact"], forces=Fp, status=sp, iters=ip)   # pinned
            # no direct pass: the list-driven class alone (horizon 20: the Riccati list class plus the dense class behind it on its second list)
            assert eng.last_host_path() == 1 and eng.launch_count() - l1 == (1 if N == 10 else 2)
            assert np.array_equal(Fp.numpy(), F) and np.array_equal(sp.numpy(), st)
            eng.close()


def test_pin_host_buffer_switches_numpy_arrays_to_zero_copy(torch_cuda):
    """plain numpy arrays take the staged path; after mpc_b200_pin_host_buffer the same arrays run zero-copy, same bits"""
    from mpc_limx_control_b200.engine import pin_host_buffer, unpin_host_buffer
    N, B, Ts = 10, 700, 0.005
    d = synth.tron1_batch(55, B, N, Ts)
    eng = make_engine(N, B, Ts=Ts)
    arrs = {k: np.ascontiguousarray(d[k]) for k in ("x0", "x_ref", "feet", "iter")}
    F0 = np.zeros((B, N, 6)); s0 = np.zeros(B, np.int32); i0 = np.zeros(B, np.int32)
    eng.solve_host(arrs["x0"], arrs["x_ref"], arrs["feet"], it=arrs["iter"], forces=F0, status=s0, iters=i0)
    assert eng.last_host_path() == 0
    F1 = np.zeros((B, N, 6)); s1 = np.zeros(B, np.int32); i1 = np.zeros(B, np.int32)
    bufs = list(arrs.values()) + [F1, s1, i1]
    for a in bufs:
        pin_host_buffer(a)
    pin_host_buffer(F1)        # pinning twice is not an error
    eng.solve_host(arrs["x0"], arrs["x_ref"], arrs["feet"], it=arrs["iter"], forces=F1, status=s1, iters=i1)
    assert eng.last_host_path() == 1
    assert np.array_equal(F0, F1) and np.array_equal(s0, s1) and np.array_equal(i0, i1) and (s1 == 0).all()
    for a in bufs:
        unpin_host_buffer(a)
    eng.solve_host(arrs["x0"], arrs["x_ref"], arrs["feet"], it=arrs["iter"], forces=F1, status=s1, iters=i1)
    assert eng.last_host_path() == 0 and np.array_equal(F0, F1)
    eng.close()


def test_device_entry_point_is_cuda_graph_capturable(torch_cuda):
    """the device entry point launches on the caller's stream and makes no synchronising call, so a control loop can
    capture several solves (here: three batches into three output buffers) in one CUDA graph and replay it"""
    torch = torch_cuda
    N, B, Ts = 10, 512, 0.005
    eng = make_engine(N, B, Ts=Ts)
    batches = [to_dev(torch, synth.tron1_batch(70 + i, B, N, Ts)) for i in range(3)]
    outs = [(torch.zeros((B, N, 6), dtype=torch.float64, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda"),
             torch.zeros(B, dtype=torch.int32, device="cuda")) for _ in range(3)]
    ref = []
    for t, o in zip(batches, outs):          # warm-up outside capture (first call configures the kernels)
        F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
        ref.append((F.clone(), st.clone()))
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for t, o in zip(batches, outs):
            eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"], forces=o[0], status=o[1], iters=o[2])
    for _ in range(3):
        for o in outs:
            o[0].zero_(); o[1].fill_(-1)
        gr.replay()
        torch.cuda.synchronize()
        for o, r in zip(outs, ref):
            assert torch.equal(o[0], r[0]) and torch.equal(o[1], r[1])
    eng.close()


def test_non_finite_instance_is_isolated(torch_cuda):
    """a NaN state poisons only its own instance: status 2 there, neighbours (same CTA) still certified"""
    torch = torch_cuda
    N, B, Ts = 10, 16, 0.005
    d = synth.tron1_batch(4, B, N, Ts)
    d["x0"][5, 9] = np.nan; d["x_ref"][5, :, 9] = np.nan
    eng = make_engine(N, B, Ts=Ts)
    t = to_dev(torch, d)
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize()
    st = st.cpu().numpy(); F = F.cpu().numpy()
    assert st[5] == 2 and (np.delete(st, 5) == 0).all()
    po = O.tron1_defaults(Ts=Ts)
    c_ref = np.stack([O.contact_schedule(int(i), N) for i in d["iter"]])
    keep = [b for b in range(B) if b != 5]
    Fo, so, _ = O.tron1_solve_batch(po, N, d["x0"][keep], d["x_ref"][keep], d["feet"][keep], c_ref[keep], nthreads=4)
    assert np.abs(F[keep] - Fo).max() / max(1.0, np.abs(Fo).max()) < 1e-4
    eng.close()


@pytest.mark.parametrize("N", [10, 20])
def test_random_contact_patterns(torch_cuda, N):
    """arbitrary contact schedules (flight phases, mixed single/double support in one horizon): compact sizes
    that are not the full 3N / 6N exercise the partial-size register paths and the capacity routing"""
    torch = torch_cuda
    B, Ts = 96, 0.01
    rng = np.random.default_rng(6)
    d = synth.tron1_batch(56, B, N, Ts)
    d["x0"][:, [0, 1, 6, 7, 8, 9, 10, 11]] *= 3
    contact = (rng.random((B, N, 2)) < rng.uniform(0.2, 0.95, (B, 1, 1))).astype(np.uint8)
    contact[0] = 0; contact[1] = 1; contact[2, :, 0] = 1; contact[2, :, 1] = 0   # all swing / all stance / single stance
    eng = make_engine(N, B, Ts=Ts)
    t = to_dev(torch, d)
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], contact=torch.from_numpy(contact).cuda())
    torch.cuda.synchronize()
    F = F.cpu().numpy(); st = st.cpu().numpy()
    assert (st == 0).all(), st
    po = O.tron1_defaults(Ts=Ts)
    Fo, so, _ = O.tron1_solve_batch(po, N, d["x0"], d["x_ref"], d["feet"], contact, nthreads=8)
    assert (so == 0).all()
    assert np.abs(F - Fo).max() / max(1.0, np.abs(Fo).max()) < 1e-4
    assert np.all(F.reshape(B, N, 2, 3)[contact == 0] == 0.0)
    nc = 3 * contact.reshape(B, -1).sum(1)
    assert ((nc > 0) & (nc < 3 * N)).any() and ((nc > 3 * N) & (nc < 6 * N)).any()   # both partial classes present
    eng.close()


def test_staged_host_path_horizon50_large_class_matches_device(torch_cuda):
    """Horizon 50 double support keeps its 364 KB factors in ONE set of global slabs.  The chunk-pipelined staged host
    path runs several chunks on different streams: their large-class kernels must not share slabs concurrently
    (they are serialised by an event).  Standing and mixed batches, >= 2 chunks, bit-identical to the device call."""
    torch = torch_cuda
    N, B, Ts = 50, 2048, 0.005
    eng = make_engine(N, B, Ts=Ts)
    for standing_frac in (1.0, 0.3):
        d = synth.tron1_batch(91, B, N, Ts)
        it = d["iter"].copy()
        it[: int(standing_frac * B)] = -1          # iter < 0: standing on both feet (n = 300)
        d["iter"] = it
        t = to_dev(torch, d)
        F, st, its = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
        torch.cuda.synchronize()
        eng.set_host_mode(eng.HOST_STAGED)
        Fh, sh, ih = eng.solve_host(d["x0"], d["x_ref"], d["feet"], it=d["iter"])
        eng.set_host_mode(eng.HOST_AUTO)
        assert eng.last_host_path() == 0
        assert np.array_equal(np.asarray(sh), st.cpu().numpy())
        assert (np.asarray(sh) == 0).all()
        assert np.array_equal(np.asarray(Fh), F.cpu().numpy())
    eng.close()


@pytest.mark.parametrize("N,B,standing_every", [(20, 1536, 5), (50, 296, 4)])
def test_every_instance_vs_oracle_horizons_20_50(torch_cuda, N, B, standing_every):
    """EVERY instance of one horizon-20 and one horizon-50 batch (trot + some standing robots: both capacity classes,
    horizon 50 = the tiled tensor-core Cholesky in shared memory AND in global slabs) against the oracle's active-set
    solution and the oracle's dense H, f -- no stride."""
    torch = torch_cuda
    Ts = 0.005
    d = synth.tron1_batch(4242 + N, B, N, Ts)
    d["iter"][::standing_every] = -1
    d["x0"][:, [0, 1, 6, 7, 8, 9, 10, 11]] *= 2.0          # some instances leave the interior face
    eng = make_engine(N, B, Ts=Ts, mu=0.4)
    t = to_dev(torch, d)
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize()
    F, st, it = F.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy()
    assert (st == 0).all()
    po = O.tron1_defaults(Ts=Ts, mu=0.4)
    c_ref = np.stack([O.contact_schedule(int(i), N) for i in d["iter"]])
    Fo, so, _ = O.tron1_solve_batch(po, N, d["x0"], d["x_ref"], d["feet"], c_ref, nthreads=os.cpu_count() or 8)
    assert (so == 0).all()
    scale = np.maximum(1.0, np.abs(Fo).reshape(B, -1).max(1))
    err = np.abs(F - Fo).reshape(B, -1).max(1) / scale
    assert err.max() < 1e-4, (int(err.argmax()), float(err.max()))
    assert np.all(F.reshape(B, N, 2, 3)[c_ref == 0] == 0.0)
    assert it.max() > 1                                        # the batch does contain multi-iteration instances
    for b in range(0, B, max(1, B // 24)):                     # KKT certificate on a sample (dense H is O(n^2 p) on the CPU)
        c = O.tron1_condense(po, N, d["x0"][b], d["x_ref"][b], d["feet"][b], want_pred=False)
        assert O.tron1_natural_residual(po, N, c["H"], c["f"], c_ref[b], F[b]) < 1e-6
    eng.close()


@pytest.mark.parametrize("N", [50, 20])
def test_horizon50_riccati_class_hands_uncertified_instances_to_the_dense_class(torch_cuda, N):
    """Horizon 50 runs its active-face solves as Riccati sweeps (one warp per instance); an instance whose active-face iteration
    does not certify within max_newton solves is appended to the overflow list and solved from scratch by the dense
    tensor-core class behind it.  max_newton = 1 forces that hand-over for every instance that leaves the interior face; the
    same batch with the default cap is solved by the Riccati class alone.  Both must agree with the oracle.
    Horizon 20: the Riccati work type is the LIST-DRIVEN class (the standing third of the batch); its hand-over goes through the
    second overflow list to a third, dense launch."""
    torch = torch_cuda
    B, Ts = 96, 0.005
    d = synth.tron1_batch(777, B, N, Ts)
    d["iter"][::3] = -1                                      # every third robot stands: two stance feet per step
    d["x0"][:, [0, 1, 6, 7, 8, 9, 10, 11]] *= 3.0            # many instances leave the interior face
    t = to_dev(torch, d)
    po = O.tron1_defaults(Ts=Ts, mu=0.3)
    c_ref = np.stack([O.contact_schedule(int(i), N) for i in d["iter"]])
    Fo, so, _ = O.tron1_solve_batch(po, N, d["x0"], d["x_ref"], d["feet"], c_ref, nthreads=os.cpu_count() or 8)
    assert (so == 0).all()
    scale = np.maximum(1.0, np.abs(Fo).reshape(B, -1).max(1))
    res = {}
    for cap in (8, 1):
        eng = make_engine(N, B, Ts=Ts, mu=0.3, max_newton=cap)
        F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
        torch.cuda.synchronize()
        F, st, it = F.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy()
        res[cap] = (F, st, it)
        assert np.isin(st, (0, 1)).all() and np.isfinite(F).all()
        ok = st == 0
        assert ok.mean() > 0.9                               # the dense class polishes its ADMM iterates on the active face
        err = np.abs(F - Fo).reshape(B, -1).max(1) / scale
        assert err[ok].max() < 1e-4, (cap, int(err.argmax()), float(err.max()))
        assert np.all(F.reshape(B, N, 2, 3)[c_ref == 0] == 0.0)
        eng.close()
    assert res[8][2].max() > 1 and (res[8][1] == 0).all()     # several faces per instance inside the Riccati class
    assert (res[1][2] > 1).any()                              # handed over: the dense class iterated (ADMM) on them


@pytest.mark.parametrize("N,B,standing_every", [(10, 3001, 0), (10, 4096, 3), (20, 2048, 0), (50, 2100, 6)])
def test_host_auto_path_large_batches(torch_cuda, N, B, standing_every):
    """Pinned buffers at full batch sizes through the host entry (AUTO and forced zero-copy), mixed capacity classes,
    horizon 50 with its shared factor slabs: bit-identical to the device call."""
    torch = torch_cuda
    from mpc_limx_control_b200.engine import Engine
    Ts = 0.005
    d = synth.tron1_batch(31 + N, B, N, Ts)
    if standing_every:
        d["iter"][::standing_every] = -1
    eng = make_engine(N, B, Ts=Ts)
    t = to_dev(torch, d)
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize()
    F, st, it = F.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy()
    pin = {k: torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory() for k in ("x0", "x_ref", "feet", "iter")}
    for mode in (Engine.HOST_AUTO, Engine.HOST_ZEROCOPY):
        Fh = torch.zeros((B, N, 6), dtype=torch.float64).pin_memory()
        sh = torch.full((B,), -7, dtype=torch.int32).pin_memory(); ih = torch.zeros(B, dtype=torch.int32).pin_memory()
        eng.set_host_mode(mode)
        eng.solve_host(pin["x0"], pin["x_ref"], pin["feet"], it=pin["iter"], forces=Fh, status=sh, iters=ih)
        assert eng.last_host_path() == 1
        assert np.array_equal(Fh.numpy(), F) and np.array_equal(sh.numpy(), st) and np.array_equal(ih.numpy(), it)
    eng.set_host_mode(Engine.HOST_AUTO)
    eng.close()


@pytest.mark.parametrize("N,B,standing_every", [(10, 1500, 5), (20, 700, 0), (50, 300, 3)])
def test_pipelined_device_entry_matches_ordered_calls(torch_cuda, N, B, standing_every):
    """mpc_b200_tron1_solve_device_pipelined: seven independent batches issued back to back on the engine's rotating
    streams (mixed capacity classes: the overflow lists and counters are per lane, horizon 50's factor slabs are
    serialised by an event), joined once -- every batch bit-identical to the stream-ordered call."""
    torch = torch_cuda
    Ts = 0.005
    eng = make_engine(N, B, Ts=Ts)
    batches, want = [], []
    for k in range(7):
        d = synth.tron1_batch(900 + k, B, N, Ts)
        if standing_every:
            d["iter"][k % standing_every::standing_every] = -1
        t = to_dev(torch, d)
        F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
        torch.cuda.synchronize()
        batches.append(t); want.append((F.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy()))
    outs = [(torch.zeros((B, N, 6), dtype=torch.float64, device="cuda"), torch.full((B,), -9, dtype=torch.int32, device="cuda"),
             torch.zeros(B, dtype=torch.int32, device="cuda")) for _ in range(7)]
    torch.cuda.synchronize()
    calls = [eng.bind_solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"], forces=o[0], status=o[1], iters=o[2], pipelined=True)
             for t, o in zip(batches, outs)]
    for rep in range(3):
        for c in calls:
            c()
    eng.join()
    torch.cuda.synchronize()
    for (F, st, it), o in zip(want, outs):
        assert np.array_equal(o[0].cpu().numpy(), F) and np.array_equal(o[1].cpu().numpy(), st) and np.array_equal(o[2].cpu().numpy(), it)
    # an ordered call after a join sees a clean engine (per-lane counters were reset by their kernels)
    t = batches[0]
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize()
    assert np.array_equal(F.cpu().numpy(), want[0][0])
    eng.close()


@pytest.mark.parametrize("N,B,standing_every", [(10, 2000, 4), (50, 260, 3)])
def test_async_host_entry_matches_device(torch_cuda, N, B, standing_every):
    """mpc_b200_tron1_solve_host_async + mpc_b200_wait: nine independent batches in pinned host arrays queued back to back
    (more than the six lanes: a lane's buffers are reused in stream order), results bit-identical to the device call;
    pageable buffers are refused."""
    torch = torch_cuda
    from mpc_limx_control_b200 import _capi
    from mpc_limx_control_b200.engine import bind_solve_host, wait
    Ts = 0.005
    eng = make_engine(N, B, Ts=Ts)
    want, calls, outs = [], [], []
    for k in range(9):
        d = synth.tron1_batch(700 + k, B, N, Ts)
        d["iter"][k % standing_every::standing_every] = -1
        t = to_dev(torch, d)
        F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
        torch.cuda.synchronize()
        want.append((F.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy()))
        pin = {k2: torch.from_numpy(np.ascontiguousarray(d[k2])).pin_memory() for k2 in ("x0", "x_ref", "feet", "iter")}
        o = (torch.zeros((B, N, 6), dtype=torch.float64).pin_memory(), torch.full((B,), -5, dtype=torch.int32).pin_memory(),
             torch.zeros(B, dtype=torch.int32).pin_memory())
        outs.append((pin, o))
        calls.append(bind_solve_host(eng, pin["x0"], pin["x_ref"], pin["feet"], it=pin["iter"], forces=o[0], status=o[1], iters=o[2],
                                     asynchronous=True))
    for rep in range(2):
        for c in calls:
            c()
        wait(eng)
        for (F, st, it), (_, o) in zip(want, outs):
            assert np.array_equal(o[0].numpy(), F) and np.array_equal(o[1].numpy(), st) and np.array_equal(o[2].numpy(), it)
            o[0].zero_(); o[1].fill_(-5)
    # pageable buffers: refused (the asynchronous entry has no staging)
    d = synth.tron1_batch(1, 8, N, Ts)
    with pytest.raises(_capi.MpcB200Error):
        bind_solve_host(eng, d["x0"], d["x_ref"], d["feet"], it=d["iter"], forces=np.zeros((8, N, 6)), status=np.zeros(8, np.int32),
                        iters=np.zeros(8, np.int32), asynchronous=True)()
    eng.close()


def test_async_control_entry_matches_synchronous(torch_cuda):
    """mpc_b200_tron1_control_host_async: controller-shaped batches queued on the lanes, first-step forces written by the
    kernels straight into pinned host arrays; bit-identical to the synchronous controller-shaped call."""
    torch = torch_cuda
    from mpc_limx_control_b200.engine import bind_control_host, wait
    N, B, Ts = 10, 1200, 0.005
    eng = make_engine(N, B, Ts=Ts)
    sets = []
    for k in range(8):
        d = synth.tron1_batch(800 + k, B, N, Ts)
        d["iter"][k % 5::5] = -1
        pin = {k2: torch.from_numpy(np.ascontiguousarray(d[k2])).pin_memory() for k2 in ("x0", "feet", "iter", "omega_yaw", "velocity_x")}
        u_s = torch.zeros((B, 6), dtype=torch.float64).pin_memory(); s_s = torch.zeros(B, dtype=torch.int32).pin_memory(); i_s = torch.zeros(B, dtype=torch.int32).pin_memory()
        bind_control_host(eng, pin["x0"], pin["omega_yaw"], pin["velocity_x"], pin["feet"], it=pin["iter"], u0=u_s, status=s_s, iters=i_s)()
        u_a = torch.zeros((B, 6), dtype=torch.float64).pin_memory(); s_a = torch.full((B,), -3, dtype=torch.int32).pin_memory(); i_a = torch.zeros(B, dtype=torch.int32).pin_memory()
        call = bind_control_host(eng, pin["x0"], pin["omega_yaw"], pin["velocity_x"], pin["feet"], it=pin["iter"], u0=u_a, status=s_a, iters=i_a,
                                 asynchronous=True)
        sets.append((pin, (u_s, s_s, i_s), (u_a, s_a, i_a), call))
    for _, _, _, call in sets:
        call()
    wait(eng)
    for _, (u_s, s_s, i_s), (u_a, s_a, i_a), _ in sets:
        assert (s_s.numpy() == 0).all()
        assert np.array_equal(u_a.numpy(), u_s.numpy()) and np.array_equal(s_a.numpy(), s_s.numpy()) and np.array_equal(i_a.numpy(), i_s.numpy())
    eng.close()


@pytest.mark.parametrize("B", [1, 7, 148])
def test_latency_class_small_standing_batches(torch_cuda, B):
    """batches no larger than the SM count send their double-support instances to the latency class of horizon 10 (8 warps
    per instance, tiled tensor-core factorisation; mpc_b200.cu: MPC_N10_WPI_LAT): every instance against the oracle's
    active set, device and host entry bit-identical, and the same instances inside a large batch (two-warp register
    elimination) agree to rounding"""
    torch = torch_cuda
    N, Ts = 10, 0.02
    big = 1024
    d = synth.tron1_batch(4242, big, N, Ts, standing=True)
    d["x0"][:, [0, 1, 6, 7, 8, 9, 10, 11]] *= 4.0                       # pyramids bind: several Newton iterations
    if B > 1:
        d["iter"][1] = 3                                                # mixed batch: one walking instance (small class)
    po = O.tron1_defaults(Ts=Ts, mu=0.3)
    eng = make_engine(N, big, Ts=Ts, mu=0.3)
    t = to_dev(torch, {k: v[:B] for k, v in d.items()})
    F, st, it = eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize()
    F = F.cpu().numpy(); st = st.cpu().numpy()
    assert (st == 0).all()
    pin = {k: torch.from_numpy(np.ascontiguousarray(d[k][:B])).pin_memory() for k in ("x0", "x_ref", "feet", "iter")}
    Fh = torch.zeros((B, N, 6), dtype=torch.float64).pin_memory()
    sh = torch.full((B,), -7, dtype=torch.int32).pin_memory(); ih = torch.zeros(B, dtype=torch.int32).pin_memory()
    eng.solve_host(pin["x0"], pin["x_ref"], pin["feet"], it=pin["iter"], forces=Fh, status=sh, iters=ih)
    assert np.array_equal(Fh.numpy(), F) and np.array_equal(sh.numpy(), st)
    tb = to_dev(torch, d)
    Fb, sb, _ = eng.solve(tb["x0"], tb["x_ref"], tb["feet"], it=tb["iter"])
    torch.cuda.synchronize()
    Fb = Fb.cpu().numpy()[:B]
    assert (sb.cpu().numpy() == 0).all()
    assert np.abs(Fb - F).max() / max(1.0, np.abs(Fb).max()) < 1e-7
    for b in range(B):
        contact = O.contact_schedule(int(d["iter"][b]), N)
        c = O.tron1_condense(po, N, d["x0"][b], d["x_ref"][b], d["feet"][b], want_pred=False)
        A, lbA, ubA, lb, ub = O.tron1_constraints(po, N, contact)
        u, info = O.qp_solve(c["H"], c["f"], A, lbA, ubA, lb, ub)
        assert info["status"] == 0
        assert np.abs(F[b].reshape(-1) - u).max() / max(1.0, np.abs(u).max()) < 1e-4
        assert O.tron1_natural_residual(po, N, c["H"], c["f"], contact, F[b]) < 1e-6
    eng.close()
