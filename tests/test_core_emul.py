"""CPU tier: the product's device source (csrc/tron1_core.cuh) compiled for the host with a
one-thread group (tests/emul) against the oracle and the golden fixtures.  Checks the kernel
MATHEMATICS without a GPU; the -m gpu tier re-runs the same comparisons through the C ABI."""
import numpy as np
import pytest

import emul_lib as E
import oracle_lib as O
from mpc_limx_control_b200 import synth


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max())


def test_golden_cases(golden):
    for key in golden["tron1_cases"]:
        N = int(golden[f"{key}_N"]); Ts = float(golden[f"{key}_Ts"]); ltv = int(golden[f"{key}_ltv"])
        p = E.default_params(Ts=Ts, ltv=ltv)
        x0 = golden[f"{key}_x0"]; xr = golden[f"{key}_xref"].T.copy(); feet = golden[f"{key}_feet"]
        c = E.dump(p, N, x0, xr, feet)
        for k in ("H", "f", "A_aug", "B_aug"):
            assert rel(c[k], golden[f"{key}_{k}"]) < 1e-9, (key, k)      # north_star: 1e-9 relative, FP64
        contact = golden[f"{key}_contact"]
        F, st, it = E.solve(p, N, x0, xr, feet, contact)
        assert st == 0, key
        assert np.abs(F.reshape(-1) - golden[f"{key}_U"]).max() / max(1.0, np.abs(F).max()) < 1e-5, key
        assert np.all(F.reshape(N, 2, 3)[contact == 0] == 0.0)


@pytest.mark.parametrize("N,Ts,ltv,standing,scale,mu", [
    (10, 0.005, 0, False, 1, 0.5), (10, 0.005, 1, False, 1, 0.5), (20, 0.005, 1, True, 1, 0.5),
    (10, 0.05, 1, False, 3, 0.5), (10, 0.001, 1, False, 1, 0.5), (10, 0.02, 1, True, 8, 0.2),
    (20, 0.05, 1, True, 6, 0.3), (4, 0.01, 1, True, 3, 0.5), (50, 0.005, 1, False, 1, 0.5), (50, 0.01, 1, True, 3, 0.4)])
def test_random_vs_oracle(N, Ts, ltv, standing, scale, mu):
    B = 10 if N < 50 else 2       # N = 50: BASELINE config 4 (n = 150 / 300 variables)
    d = synth.tron1_batch(31, B, N, Ts, standing=standing)
    po = O.tron1_defaults(Ts=Ts, ltv=ltv, mu=mu); pe = E.default_params(Ts=Ts, ltv=ltv, mu=mu)
    for b in range(B):
        x0 = d["x0"][b].copy(); x0[[0, 1, 6, 7, 8, 9, 10, 11]] *= scale
        xr = d["x_ref"][b]; feet = d["feet"][b]
        c = O.tron1_condense(po, N, x0, xr, feet); c2 = E.dump(pe, N, x0, xr, feet)
        for k in ("H", "f", "A_aug", "B_aug"):
            assert rel(c2[k], c[k]) < 1e-9, k
        contact = O.contact_schedule(int(d["iter"][b]), N)
        assert np.array_equal(contact, E.gait_contact(pe, int(d["iter"][b]), N))     # bit-exact mode indices
        A, lbA, ubA, lb, ub = O.tron1_constraints(po, N, contact)
        u, info = O.qp_solve(c["H"], c["f"], A, lbA, ubA, lb, ub)
        F, st, it = E.solve(pe, N, x0, xr, feet, contact)
        assert st == 0 and info["status"] == 0
        assert np.abs(F.reshape(-1) - u).max() / max(1.0, np.abs(u).max()) < 1e-4   # north_star tolerance
        assert O.tron1_natural_residual(po, N, c["H"], c["f"], contact, F) < 1e-6   # KKT residual bound


def test_admm_fallback_path_tiled_horizon50():
    """Horizon 50 stores the reduced Hessian as 8x8 tiles: max_newton=1 sends instances through the ADMM, i.e. through the
    blocked forward AND backward solves on that layout (the face solve alone only needs the backward one)."""
    N, Ts = 50, 0.02
    d = synth.tron1_batch(12, 2, N, Ts, standing=False)
    po = O.tron1_defaults(Ts=Ts, mu=0.3); pe = E.default_params(Ts=Ts, mu=0.3, max_newton=1)
    used_admm = 0
    for b in range(2):
        x0 = d["x0"][b].copy(); x0[[0, 1, 6, 7, 8, 9, 10, 11]] *= 6
        contact = O.contact_schedule(int(d["iter"][b]), N)
        c = O.tron1_condense(po, N, x0, d["x_ref"][b], d["feet"][b], want_pred=False)
        A, lbA, ubA, lb, ub = O.tron1_constraints(po, N, contact)
        u, info = O.qp_solve(c["H"], c["f"], A, lbA, ubA, lb, ub)
        F, st, it = E.solve(pe, N, x0, d["x_ref"][b], d["feet"][b], contact)
        used_admm += it > 1
        assert st == 0 and info["status"] == 0
        assert np.abs(F.reshape(-1) - u).max() / max(1.0, np.abs(u).max()) < 1e-4
        assert O.tron1_natural_residual(po, N, c["H"], c["f"], contact, F) < 1e-6
    assert used_admm >= 1


@pytest.mark.parametrize("max_newton,scale,mu", [(12, 1, 0.5), (12, 8, 0.2), (1, 8, 0.2)])
def test_latency_class_tiled_storage_horizon10(max_newton, scale, mu):
    """Small batches of double-support instances of horizon 10 run in the tiled 8x8 storage (mpc_b200.cu: MPC_N10_WPI_LAT):
    same forces as the packed-storage path and as the oracle's active set, on the Newton path and (max_newton=1) the ADMM."""
    N, Ts = 10, 0.02
    d = synth.tron1_batch(77, 8, N, Ts, standing=True)
    po = O.tron1_defaults(Ts=Ts, mu=mu); pe = E.default_params(Ts=Ts, mu=mu, max_newton=max_newton)
    rng = np.random.default_rng(5)
    for b in range(8):
        x0 = d["x0"][b].copy(); x0[[0, 1, 6, 7, 8, 9, 10, 11]] *= scale
        contact = np.ones((N, 2), np.uint8)
        if b >= 4:                                   # ragged: some foot-steps in swing (nc < 60: padded tiles)
            contact[rng.integers(0, N, 3), rng.integers(0, 2, 3)] = 0
        c = O.tron1_condense(po, N, x0, d["x_ref"][b], d["feet"][b], want_pred=False)
        A, lbA, ubA, lb, ub = O.tron1_constraints(po, N, contact)
        u, info = O.qp_solve(c["H"], c["f"], A, lbA, ubA, lb, ub)
        F, st, it = E.solve_tiled60(pe, x0, d["x_ref"][b], d["feet"][b], contact)
        assert st == 0 and info["status"] == 0
        assert np.abs(F.reshape(-1) - u).max() / max(1.0, np.abs(u).max()) < 1e-4
        assert O.tron1_natural_residual(po, N, c["H"], c["f"], contact, F) < 1e-6
        assert np.all(F.reshape(N, 2, 3)[contact == 0] == 0.0)


@pytest.mark.parametrize("N,Ts,standing,scale,mu", [(10, 0.005, False, 1, 0.5), (10, 0.02, True, 8, 0.2), (20, 0.05, True, 6, 0.3),
                                                    (4, 0.01, True, 3, 0.5), (50, 0.01, True, 3, 0.4)])
def test_riccati_face_solves_match_dense_path_and_oracle(N, Ts, standing, scale, mu):
    """The Riccati work type (tron1_core.cuh: riccati_face_solve; opt-in on the device, -DMPC_RIC_N*) replaces the factorisation
    of the condensed Hessian by one backward and one forward sweep over the horizon.  Same faces, same iteration counts and the
    same forces as the dense path (to rounding), and the oracle's active-set solution within the north_star tolerance."""
    B = 4 if N < 50 else 1
    d = synth.tron1_batch(31, B, N, Ts, standing=standing)
    po = O.tron1_defaults(Ts=Ts, mu=mu); pe = E.default_params(Ts=Ts, mu=mu)
    rng = np.random.default_rng(N)
    for b in range(B):
        x0 = d["x0"][b].copy(); x0[[0, 1, 6, 7, 8, 9, 10, 11]] *= scale
        contact = O.contact_schedule(int(d["iter"][b]), N)
        if b == 1:                                   # ragged pattern: flight steps and single support mixed in
            contact = contact.copy(); contact[rng.integers(0, N, 3), rng.integers(0, 2, 3)] = 0
        F, st, it, deferred = E.solve_riccati(pe, N, x0, d["x_ref"][b], d["feet"][b], contact)
        Fx, stx, itx, dfx = E.solve_riccati(pe, N, x0, d["x_ref"][b], d["feet"][b], contact, ext_gains=True)   # gains in external storage
        assert np.array_equal(F, Fx) and (stx, itx, dfx) == (st, it, deferred)
        F2, st2, it2 = E.solve(pe, N, x0, d["x_ref"][b], d["feet"][b], contact)
        assert st == 0 and st2 == 0 and not deferred and it == it2
        assert np.abs(F - F2).max() / max(1.0, np.abs(F2).max()) < 1e-8
        c = O.tron1_condense(po, N, x0, d["x_ref"][b], d["feet"][b], want_pred=False)
        A, lbA, ubA, lb, ub = O.tron1_constraints(po, N, contact)
        u, info = O.qp_solve(c["H"], c["f"], A, lbA, ubA, lb, ub)
        assert info["status"] == 0
        assert np.abs(F.reshape(-1) - u).max() / max(1.0, np.abs(u).max()) < 1e-4
        assert O.tron1_natural_residual(po, N, c["H"], c["f"], contact, F) < 1e-6
        assert np.all(F.reshape(N, 2, 3)[contact == 0] == 0.0)


@pytest.mark.parametrize("N,ltv,seed", [(4, 1, 0), (10, 1, 1), (10, 0, 2), (20, 1, 3), (20, 0, 4)])
def test_riccati_random_contact_patterns_match_dense_path(N, ltv, seed):
    """Arbitrary contact patterns (flight steps, single and double support mixed, a foot that never lands), the frozen-model
    and the per-step-model modes, scaled states that leave the interior face: the Riccati sweeps (gains inside the struct and
    in external storage) reproduce the dense path's faces, iteration counts and forces."""
    Ts, B = 0.01, 6
    rng = np.random.default_rng(100 + seed)
    d = synth.tron1_batch(50 + seed, B, N, Ts, standing=True)
    pe = E.default_params(Ts=Ts, ltv=ltv, mu=0.35)
    po = O.tron1_defaults(Ts=Ts, ltv=ltv, mu=0.35)
    n_deferred = 0
    for b in range(B):
        x0 = d["x0"][b].copy(); x0[[0, 1, 6, 7, 8, 9, 10, 11]] *= (1.0 + b)
        contact = (rng.random((N, 2)) < 0.7).astype(np.uint8)
        if b == 0: contact[:, 1] = 0                     # the right foot never lands
        if b == 1: contact[N // 2] = 0                   # a flight step in the middle of the horizon
        if contact.sum() == 0: contact[0, 0] = 1
        F2, st2, it2 = E.solve(pe, N, x0, d["x_ref"][b], d["feet"][b], contact)
        for ext in (False, True):
            F, st, it, deferred = E.solve_riccati(pe, N, x0, d["x_ref"][b], d["feet"][b], contact, ext_gains=ext)
            # an instance the active-face iteration does not certify is handed to the dense path (which then pays its own
            # iterations on top of the max_newton sweeps already spent)
            assert st == 0 and st2 == 0 and it == it2 + (pe.max_newton if deferred else 0), (b, ext, st, st2, it, it2, deferred)
            n_deferred += int(bool(deferred))
            assert np.abs(F - F2).max() / max(1.0, np.abs(F2).max()) < 1e-8
            assert np.all(F.reshape(N, 2, 3)[contact == 0] == 0.0)
        c = O.tron1_condense(po, N, x0, d["x_ref"][b], d["feet"][b], want_pred=False)
        assert O.tron1_natural_residual(po, N, c["H"], c["f"], contact, F) < 1e-6
    assert n_deferred <= 2 * 2                              # the hand-over is the exception


def test_admm_fallback_path():
    """max_newton=1 forces every instance whose first face guess is wrong through ADMM + polish."""
    N, Ts, B = 10, 0.02, 12
    d = synth.tron1_batch(9, B, N, Ts, standing=True)
    po = O.tron1_defaults(Ts=Ts); pe = E.default_params(Ts=Ts, max_newton=1)
    used_admm = 0
    for b in range(B):
        x0 = d["x0"][b].copy(); x0[[0, 1, 6, 7, 8, 9, 10, 11]] *= 5
        contact = np.ones((N, 2), np.uint8)
        c = O.tron1_condense(po, N, x0, d["x_ref"][b], d["feet"][b], want_pred=False)
        A, lbA, ubA, lb, ub = O.tron1_constraints(po, N, contact)
        u, info = O.qp_solve(c["H"], c["f"], A, lbA, ubA, lb, ub)
        F, st, it = E.solve(pe, N, x0, d["x_ref"][b], d["feet"][b], contact)
        used_admm += it > 1
        assert st == 0
        assert np.abs(F.reshape(-1) - u).max() / max(1.0, np.abs(u).max()) < 1e-6
    assert used_admm > 0


def test_iteration_cap_reports_status():
    N, Ts = 10, 0.02
    d = synth.tron1_batch(9, 6, N, Ts, standing=True)
    pe = E.default_params(Ts=Ts, max_newton=1, max_admm=2)
    sts = []
    for b in range(6):
        x0 = d["x0"][b].copy(); x0[[0, 1, 6, 7, 8, 9, 10, 11]] *= 5
        F, st, it = E.solve(pe, N, x0, d["x_ref"][b], d["feet"][b], np.ones((N, 2), np.uint8))
        sts.append(st)
        assert np.isfinite(F).all()
        # the returned iterate is always feasible
        F3 = F.reshape(N, 2, 3)
        assert (np.abs(F3[..., 0]) <= 0.5 * F3[..., 2] + 1e-9).all() and (F3[..., 2] >= -1e-12).all()
    assert 1 in sts   # ST_MAXITER surfaced, never silently "solved"


def test_all_swing_gives_zero():
    N = 10
    pe = E.default_params()
    d = synth.tron1_batch(1, 1, N, 0.005)
    F, st, it = E.solve(pe, N, d["x0"][0], d["x_ref"][0], d["feet"][0], np.zeros((N, 2), np.uint8))
    assert st == 0 and np.all(F == 0.0)


def test_gait_matches_golden(golden):
    pe = E.default_params()
    for it, l, r in zip(golden["gait_iter"][::7], golden["gait_left"][::7], golden["gait_right"][::7]):
        c = E.gait_contact(pe, int(it), 1)
        assert c[0, 0] == 1 - l and c[0, 1] == 1 - r
    assert np.all(E.gait_contact(pe, -1, 5) == 1)   # standing


def test_gait_exact_remainder_vs_library_fmod():
    """gait_contact evaluates fmod((double)(iter*dt), (double)(swing+stance)) by an exact quotient/remainder
    instead of the library routine: compare with the oracle (C fmod) on random and adversarial loop counters,
    for the reference timing and for cycle lengths that are not powers of two (inexact reciprocal)."""
    rng = np.random.default_rng(12)
    its = np.concatenate([rng.integers(0, 2**31 - 64, 4000), np.arange(0, 3000), np.arange(10_773_900, 10_774_100),
                          500 * np.arange(1, 2000), 500 * np.arange(1, 2000) - 1, [2**31 - 64, 16_777_216, 16_777_217]])
    for dt, sw, stc in ((0.001, 0.5, 0.5), (0.001, 0.3, 0.4), (0.002, 0.35, 0.15), (0.0005, 0.7, 0.1)):
        pe = E.default_params(gait_dt=dt, gait_swing_time=sw, gait_stance_time=stc)
        g = O.gait_defaults(); g.dt = dt; g.swing_time = sw; g.stance_time = stc
        for it in its:
            a = E.gait_contact(pe, int(it), 1)
            b = O.contact_schedule(int(it), 1, g)
            assert np.array_equal(a, b), (dt, sw, stc, int(it))


def test_rollout_vs_oracle():
    """closed loop (BASELINE configs[4] shape, short): emulated device source vs the oracle loop"""
    N, Ts, steps = 10, 0.005, 40
    d = synth.tron1_batch(1004, 4, N, Ts)
    pe = E.default_params(Ts=Ts); po = O.tron1_defaults(Ts=Ts)
    offl = list(pe.foot_offset_left); offr = list(pe.foot_offset_right)
    for b, it0 in zip(range(4), (int(d["iter"][0]), 480, -1, int(d["iter"][3]))):   # 480: the gait switches mid-rollout
        xe, Ue, bad_e, its = E.rollout(pe, N, steps, d["x0"][b], d["omega_yaw"][b], d["velocity_x"][b], it0)
        xo, Uo, bad_o = O.tron1_rollout(po, N, steps, d["x0"][b], d["omega_yaw"][b], d["velocity_x"][b], it0, offl, offr)
        assert bad_e == 0 and bad_o == 0
        assert np.abs(Ue - Uo).max() / max(1.0, np.abs(Uo).max()) < 1e-6
        assert np.abs(xe - xo).max() < 1e-8
        assert its >= steps


def test_non_finite_input_reports_failure():
    """NaN / inf in the state must surface as status 2 (never a silent 'solved'), SURVEY.md 8b error convention"""
    N = 10
    pe = E.default_params()
    d = synth.tron1_batch(2, 2, N, 0.005)
    c = O.contact_schedule(int(d["iter"][0]), N)
    for bad in (np.nan, np.inf):
        x0 = d["x0"][0].copy(); x0[9] = bad
        xr = d["x_ref"][0].copy(); xr[:, 9] = bad
        F, st, it = E.solve(pe, N, x0, xr, d["feet"][0], c)
        assert st == 2


def test_random_contact_patterns():
    """arbitrary schedules: flight phases, mixed single/double support inside one horizon (n_c not 30/60)"""
    N, Ts = 10, 0.01
    rng = np.random.default_rng(5)
    d = synth.tron1_batch(55, 12, N, Ts)
    d["x0"][:, [0, 1, 6, 7, 8, 9, 10, 11]] *= 3
    po = O.tron1_defaults(Ts=Ts); pe = E.default_params(Ts=Ts)
    for b in range(12):
        contact = (rng.random((N, 2)) < (0.3 + 0.05 * b)).astype(np.uint8)
        c = O.tron1_condense(po, N, d["x0"][b], d["x_ref"][b], d["feet"][b], want_pred=False)
        A, lbA, ubA, lb, ub = O.tron1_constraints(po, N, contact)
        u, info = O.qp_solve(c["H"], c["f"], A, lbA, ubA, lb, ub)
        F, st, it = E.solve(pe, N, d["x0"][b], d["x_ref"][b], d["feet"][b], contact)
        assert st == 0 and info["status"] == 0
        assert np.abs(F.reshape(-1) - u).max() / max(1.0, np.abs(u).max()) < 1e-6
        assert np.all(F.reshape(N, 2, 3)[contact == 0] == 0.0)
