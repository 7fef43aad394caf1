#!/usr/bin/env python
"""bench.py -- batched MPC solves/sec (TRON1) on N B200s + p50 single-solve latency.

One "step" = one pass of the hot path (linearise -> discretise -> condense -> QP solve -> forces) over one batch of
synthetic instances.  Instances are independent, so GPUs shard by instance with no collective on the solve path
(torch.distributed is used only for the barrier and the max-over-ranks time).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|2s|2x|3|4|5] [--impl reference]

--config selects a BASELINE.json workload (default 2 = the configuration the headline metric is quoted on):
  2   B=4096 per GPU, horizon 10, trot schedule from the gait clock, friction pyramid   (weak scaling)
  2s  same, every robot standing on both feet (double-support capacity class)
  2x  same, stressed: state x5, mu=0.3, Ts=0.02 (friction pyramid active, multi-iteration instances)
  3   B=65536 TOTAL, horizon 20, sharded 65536/G per GPU                                 (strong scaling)
  4   B=8192 TOTAL, horizon 50 (tensor-core Cholesky path), sharded                      (strong scaling)
  5   closed loop: 16384 robots TOTAL x 1000 control steps, warm-started, sharded        (strong scaling; a step =
      one full rollout; --steps is clamped so the default run stays within minutes)
--impl reference times the reference's CPU path (the oracle port of QPSolver/mpcQP + a cold-start active-set QP; Eigen /
qpOASES are not installable here) on all host cores, on a bounded sample of the same workload."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly one JSON line.  Libraries write there too (under torchrun NCCL prints a version banner on
# file descriptor 1), so the real stdout is kept aside for the JSON line and descriptor 1 is pointed at stderr.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


CONFIGS = {
    "2": dict(N=10, B=4096, seed=1001, Ts=0.005, standing=False, scale=1.0, mu=0.5, kind="solve", scaling="weak",
              what="trot contact schedule from the gait clock, friction pyramid mu=0.5, f_max=2mg (BASELINE configs[1])"),
    "2s": dict(N=10, B=4096, seed=1001, Ts=0.005, standing=True, scale=1.0, mu=0.5, kind="solve", scaling="weak",
               what="every robot standing on both feet (double support), friction pyramid mu=0.5 (BASELINE configs[0] batched)"),
    "2x": dict(N=10, B=4096, seed=1001, Ts=0.02, standing=False, scale=5.0, mu=0.3, kind="solve", scaling="weak",
               what="stressed: state x5, mu=0.3, Ts=0.02, trot schedule (friction pyramid active)"),
    "3": dict(N=20, B=65536, seed=1002, Ts=0.005, standing=False, scale=1.0, mu=0.5, kind="solve", scaling="strong",
              what="randomised yaw / foot positions / gait phase, sharded by instance (BASELINE configs[2])"),
    "4": dict(N=50, B=8192, seed=1003, Ts=0.005, standing=False, scale=1.0, mu=0.5, kind="solve", scaling="strong",
              what="long horizon, trot gait (150 of the 300 condensed variables active per instance), active-face solves as Riccati sweeps; "
                   "the tiled FP64 tensor-core Cholesky class runs behind it for uncertified instances (BASELINE configs[3])"),
    "5": dict(N=10, B=16384, seed=1004, Ts=0.005, standing=False, scale=1.0, mu=0.5, kind="rollout", steps=1000, scaling="strong",
              what="closed loop: linearise -> condense -> solve -> integrate, warm-started, state resident on the GPU (BASELINE configs[4])"),
}


def metric_name(cfg):
    return f"batched MPC solves/sec (TRON1, N={cfg['N']})"


def workload_name(cfg, world):
    """The SAME string in both arms (the driver compares config.workload of the two lines)."""
    per = "per GPU" if cfg["scaling"] == "weak" else f"in total, sharded over {world} GPU(s)"
    s = f"TRON1 convex MPC, horizon {cfg['N']}, Ts {cfg['Ts']}, B={cfg['B']} instances {per}"
    if cfg["kind"] == "rollout":
        s += f" x {cfg['steps']} closed-loop control steps"
    return s + ", " + cfg["what"]


def dense_equiv_flops(N, iters):
    """SURVEY.md 8d DENSE accounting (n = 6N, p = 13(N+1)): what a dense evaluation of the reference's formulas costs."""
    n, p = 6 * N, 13 * (N + 1)
    return 6422 * N + (n * n * p + n * p) + (26 * p + 2 * p + 2 * p * n) + n ** 3 / 3 + iters * (4 * n * n + 10 * n + 32 * N)


def structured_flops(N, nc, iters, group_threads):
    """FP64 operations the kernel EXECUTES per solve (thread level, FMA = 2), hand-counted from csrc/tron1_core.cuh
    (DESIGN.md section 4 'executed work').  m = nc/3 stance foot-steps.
      setup (once)            : model + horizon sums + free response + adjoint        ~ 450 N
      per active-face solve   : m(m+1)/2 Hessian blocks x 157  (build_hessian, one 3x3 block per work item)
                                + elimination / factorisation (below)
                                + gradient + optimality check                         ~ 300 N
      elimination, nc <= 60   : register Gauss-Jordan, EVERY lane of the group carries a row window:
                                lanes x sum over the 5 stages of (nc/5) columns x (2 W + 10), W = window length
      horizon 50              : Riccati class (see the branch below); the dense tiled DMMA Cholesky class only takes the
                                instances whose active-face iteration does not certify (none in the bench workloads)
    Returns (total, tensor_part).  Checked against ncu's executed thread-level DFMA/DMUL/DADD count (which excludes DMMA)
    on the committed captures profiles/r2_c*: config 2 0.999, 2s 1.06, 3 1.01, 4 (non-tensor part) within 10 %."""
    m = nc / 3.0
    if N == 50:
        # Riccati class (horizon 50): no matrix, one backward and one forward sweep per active-face solve.  Per step with nf stance
        # feet (counted per ACTIVE lane of the warp, phases A-E of riccati_backward_step): nf = 0: 888, nf = 1: 3373, nf = 2: 7595
        # (A: 360 + 372 nf, B: 468 + 108 nf, C: 237 / 690, D: 13 columns x (3x3-block cofactor inverses + solves + force-space
        # rows) = 832 / 3185, E: 12 rows x (26 mm + 5)); forward step 12 lanes x 53 + 81 nf; gradient (one adjoint pass) + check
        # ~ 220 N; setup without f ~ 300 N.  Checked against ncu: 233.6 k computed against 246.5 k executed per trot solve
        # (profiles/r2e_c4_solve_kernel.json), 413 k against 410 k per double-support solve with the first forward step version.
        nf = m / N                                   # stance feet per step (1 trot, 2 double support)
        back = N * (888.0 + (3373.0 - 888.0) * min(nf, 1.0) + (7595.0 - 3373.0) * max(nf - 1.0, 0.0))
        fwd = N * (636.0 + 81.0 * nf)
        return 300.0 * N + iters * (back + fwd + 220.0 * N), 0.0
    setup = 450.0 * N
    hess = 157.0 * m * (m + 1) / 2
    it_rest = 300.0 * N
    tensor = 0.0
    lanes = group_threads
    elim = lanes * sum((nc / 5.0) * (2 * (nc - s * nc / 5.0) + 10) for s in range(5))
    return setup + iters * (hess + elim + it_rest), iters * tensor


def algorithmic_bytes(N):
    return 104 + 104 * (N + 1) + 48 + 4 + 48 * N + 8


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        out["samples"] = len(rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, nme in enumerate(names):
            if any("Active" == r[5 + i].strip() for r in rows):
                out["reasons"].append(nme)
        return out


def synth_batch(cfg, B, first=0):
    from mpc_limx_control_b200 import synth
    d = synth.tron1_batch(cfg["seed"], B, cfg["N"], cfg["Ts"], first=first, standing=cfg["standing"])
    if cfg["scale"] != 1.0:
        d["x0"][:, [0, 1, 6, 7, 8, 9, 10, 11]] *= cfg["scale"]
    return d


# ------------------------------------------------------------------------------------------------------------------
# CPU path: the oracle port on the host cores (cpu_baseline of the GPU line, and the whole --impl reference arm)
class CpuPath:
    def __init__(self, cfg, nthreads):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        self.O, self.cfg, self.nthreads = O, cfg, nthreads
        self.p = O.tron1_defaults(Ts=cfg["Ts"], mu=cfg["mu"])

    def prepare(self, n):
        cfg, O = self.cfg, self.O
        self.d = synth_batch(cfg, n)
        self.c = np.stack([O.contact_schedule(int(i), cfg["N"]) for i in self.d["iter"]])
        self.n = n

    def run(self, rollout_steps=None):
        """one pass over the prepared sample; returns the number of solves done"""
        cfg, O, d = self.cfg, self.O, self.d
        if cfg["kind"] == "rollout":
            from concurrent.futures import ThreadPoolExecutor
            from mpc_limx_control_b200 import synth
            steps = rollout_steps

            def one(b):
                O.tron1_rollout(self.p, cfg["N"], steps, d["x0"][b], float(d["omega_yaw"][b]), float(d["velocity_x"][b]),
                                int(d["iter"][b]), synth.FOOT_OFFSET_L, synth.FOOT_OFFSET_R)
            with ThreadPoolExecutor(self.nthreads) as ex:      # ctypes releases the GIL: one rollout per thread at a time
                list(ex.map(one, range(self.n)))
            return self.n * steps
        O.tron1_solve_batch(self.p, cfg["N"], d["x0"], d["x_ref"], d["feet"], self.c, self.nthreads)
        return self.n

    def size_for(self, seconds, cap, rollout_steps=None):
        """sample size that takes about `seconds` on this box"""
        probe = max(self.nthreads * (1 if self.cfg["N"] >= 50 or self.cfg["kind"] == "rollout" else 4), 8)
        self.prepare(probe)
        self.run(rollout_steps)     # warm-up
        t = time.perf_counter()
        done = self.run(rollout_steps)
        rate = done / (time.perf_counter() - t)
        per_item = done / probe
        n = int(min(cap, max(self.nthreads, rate * seconds / per_item)))
        self.prepare(n)
        return n


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cores = os.cpu_count() or 1
    cpu = CpuPath(cfg, cores)
    rsteps = 20 if cfg["kind"] == "rollout" else None
    # bounded sample per step so that the whole K+W run ends within a few minutes
    n = cpu.size_for(min(2.0, 150.0 / (args.steps + args.warmup)), cfg["B"], rsteps)
    for _ in range(args.warmup):
        cpu.run(rsteps)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        done += cpu.run(rsteps)
    dt = time.perf_counter() - t0
    v = done / dt
    sample = (f"{n} of the {cfg['B']} instances per step (seed {cfg['seed']})"
              + (f", first {rsteps} of the {cfg['steps']} control steps" if rsteps else "")
              + f", {cores} threads, one solve per thread at a time")
    emit({
        "impl": "reference", "metric": metric_name(cfg), "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg, world), "seed": cfg["seed"], "config_id": args.config, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "restated reference CPU path (Eigen/qpOASES unavailable): dense condensing + cold-start active set"},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------------------------
def run_gpu(args, cfg):
    import torch
    import torch.distributed as dist
    from mpc_limx_control_b200 import shard
    from mpc_limx_control_b200.engine import Engine, measure_fp64_peak, bind_solve_host, bind_control_host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local)
    N, K, W = cfg["N"], args.steps, args.warmup
    rollout = cfg["kind"] == "rollout"
    # per-rank share: weak scaling keeps B per GPU, strong scaling block-partitions the total (shard.partition)
    if cfg["scaling"] == "weak":
        B, first_inst, B_total = cfg["B"], rank * cfg["B"], world * cfg["B"]
    else:
        first_inst, B = shard.partition(cfg["B"], world, rank)
        B_total = cfg["B"]
    if rollout:
        K = max(1, min(K, 3)); W = max(1, min(W, 1))      # one step = 1000 control steps of every robot (~0.13 s)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = Engine(horizon=N, max_batch=B, device=local, Ts=cfg["Ts"], mu=cfg["mu"])
    per_batch_in = B * (104 + 104 * (N + 1) + 48 + 4)
    extra = {}

    if not rollout:
        # rotating pool of distinct input batches, larger than the 126 MB L2, resident in HBM
        pool_n = max(2, int(np.ceil(160e6 / per_batch_in)))
        pool = []
        for i in range(pool_n):
            d = synth_batch(cfg, B, first=first_inst + i * B_total)
            pool.append({k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "x_ref", "feet", "iter")})
        forces = torch.empty((B, N, 6), dtype=torch.float64, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        iters = torch.empty(B, dtype=torch.int32, device=dev)
        # one pre-bound C-ABI call per pool entry: the timed loop issues mpc_b200_tron1_solve_device and nothing else
        calls = [eng.bind_solve(p["x0"], p["x_ref"], p["feet"], it=p["iter"], forces=forces, status=status, iters=iters) for p in pool]
        # the timed loop works through a queue of independent batches: mpc_b200_tron1_solve_device_pipelined lets the CTAs of
        # batch k+1 fill the slots the last multi-iteration instances of batch k leave idle; one join before the closing event.
        # --serialized times the stream-ordered entry instead (every batch waits for the previous one to drain).
        pcalls = [eng.bind_solve(p["x0"], p["x_ref"], p["feet"], it=p["iter"], forces=forces, status=status, iters=iters, pipelined=True)
                  for p in pool]
        use = calls if args.serialized else pcalls
        step = lambda i: use[i % pool_n]()
        l2_note = f"rotating pool of {pool_n} distinct input batches ({pool_n * per_batch_in / 1e6:.0f} MB) > 126 MB L2"
    else:
        d = synth_batch(cfg, B, first=first_inst)
        x_init = torch.from_numpy(d["x0"]).to(dev)
        oy = torch.from_numpy(d["omega_yaw"]).to(dev); vx = torch.from_numpy(d["velocity_x"]).to(dev)
        it0 = torch.from_numpy(d["iter"]).to(dev)
        x_state = x_init.clone()
        res = {}

        def step(i):
            x_state.copy_(x_init)
            _, res["bad"], res["its"] = eng.rollout(x_state, oy, vx, it0, cfg["steps"])
        l2_note = "closed loop: the state never leaves the SM between control steps; inputs are 228 B per robot"

    for i in range(W):
        step(i)
    eng.join()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        step(W + i)
    eng.join()          # the current stream waits for every pipelined batch: e1 is behind all K steps
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    ms_serial = None
    if not rollout and not args.serialized:
        # the same K steps through the stream-ordered entry (one batch at a time): the latency-bound figure
        for i in range(min(W, 10)):
            calls[i % pool_n]()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Ks = min(K, 500)
        s0.record()
        for i in range(Ks):
            calls[(W + i) % pool_n]()
        s1.record()
        torch.cuda.synchronize()
        ms_serial = s0.elapsed_time(s1) / Ks
    ms_max = max_over_ranks(ms)
    if rollout:
        solves_per_step = B_total * cfg["steps"]
        n_bad = int(res["bad"].sum().item())
        mean_iters = float(res["its"].float().mean().item()) / cfg["steps"]
    else:
        solves_per_step = B_total
        n_bad = int((status != 0).sum().item())
        mean_iters = float(iters.float().mean().item())
    value = solves_per_step * K / (ms_max * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside) -------
    e2e, e2e_ctrl = None, None
    if not rollout:
        Ke = min(500, max(5, K // 4)) if N == 10 else max(3, min(10, K // 2))
        host_calls, keep = [], []
        for j in range(4 if N == 10 else 2):
            dj = synth_batch(cfg, B, first=first_inst + (j + 40) * B_total)
            pj = {k: torch.from_numpy(dj[k]).pin_memory() for k in ("x0", "x_ref", "feet", "iter")}
            out = (torch.empty((B, N, 6), dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory(),
                   torch.empty(B, dtype=torch.int32).pin_memory())
            keep.append((dj, pj, out))
            host_calls.append(bind_solve_host(eng, pj["x0"], pj["x_ref"], pj["feet"], it=pj["iter"], forces=out[0], status=out[1], iters=out[2]))
        nh = len(host_calls)
        for j in range(max(3, min(W, 8))):
            host_calls[j % nh]()
        barrier()
        t0 = time.perf_counter()
        for j in range(Ke):
            host_calls[j % nh]()      # one mpc_b200_tron1_solve_host: inputs over PCIe, solve, results back, sync
        torch.cuda.synchronize()
        te = max_over_ranks(time.perf_counter() - t0)
        sync_value = B_total * Ke / te
        sync_path = "zero-copy (kernel reads/writes the pinned host buffers over PCIe)" if eng.last_host_path() else "staged copies"
        # the same batches through the asynchronous host entry: queue Ke independent batches (pinned host inputs read by the
        # kernels over PCIe, results through per-lane device buffers + copy engine into the pinned host outputs), wait once.
        # Every input byte crosses PCIe and every result byte is back in host memory inside the timed region.
        from mpc_limx_control_b200.engine import wait as eng_wait
        acalls = [bind_solve_host(eng, pj["x0"], pj["x_ref"], pj["feet"], it=pj["iter"], forces=out[0], status=out[1], iters=out[2],
                                  asynchronous=True) for (dj, pj, out) in keep]
        for j in range(max(3, min(W, 8))):
            acalls[j % nh]()
        eng_wait(eng)
        barrier()
        t0 = time.perf_counter()
        for j in range(Ke):
            acalls[j % nh]()
        eng_wait(eng)
        te = max_over_ranks(time.perf_counter() - t0)
        async_value = B_total * Ke / te
        async_path = ("mpc_b200_tron1_solve_host_async x steps + mpc_b200_wait: pinned host inputs read by the kernels over PCIe, results "
                      "through device buffers + copy engine into pinned host outputs, consecutive batches overlapped")
        # both are public entry points; the headline is the one a caller with this batch shape would use (the better one)
        e2e = {"value": max(async_value, sync_value), "unit": "solves/s", "h2d_bytes_per_step": B_total * (104 + 104 * (N + 1) + 48 + 4),
               "d2h_bytes_per_step": B_total * (48 * N + 8), "steps": Ke,
               "path": async_path if async_value >= sync_value else "mpc_b200_tron1_solve_host per step: " + sync_path,
               "host_batches": f"{nh} distinct pinned input/output batches in rotation",
               "asynchronous_entry_value": async_value, "synchronous_entry_value": sync_value, "synchronous_entry_path": sync_path}
        if N == 10:
            # the same call forced onto the staged-copy path (what a caller with pageable buffers gets), for comparison
            eng.set_host_mode(Engine.HOST_STAGED)
            for _ in range(3):
                host_calls[0]()
            barrier()
            ns = max(5, Ke // 4)
            t0 = time.perf_counter()
            for _ in range(ns):
                host_calls[0]()
            torch.cuda.synchronize()
            e2e["staged_copies_value"] = B_total * ns / max_over_ranks(time.perf_counter() - t0)
            eng.set_host_mode(Engine.HOST_AUTO)
            # controller-shaped host call (command in, first-step force out: the reference mpcQP's own I/O)
            dj, pj, out = keep[0]
            pin_c = {k: torch.from_numpy(dj[k]).pin_memory() for k in ("omega_yaw", "velocity_x")}
            u0h = torch.empty((B, 6), dtype=torch.float64).pin_memory()
            ctrl_call = bind_control_host(eng, pj["x0"], pin_c["omega_yaw"], pin_c["velocity_x"], pj["feet"], it=pj["iter"], u0=u0h,
                                          status=out[1], iters=out[2])
            for _ in range(3):
                ctrl_call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(Ke):
                ctrl_call()
            torch.cuda.synchronize()
            ctrl_sync = B_total * Ke / max_over_ranks(time.perf_counter() - t0)
            # the same through the asynchronous entry: Ke independent batches queued, one wait
            u0s = [torch.empty((B, 6), dtype=torch.float64).pin_memory() for _ in range(nh)]
            pcs = [{k: torch.from_numpy(dj2[k]).pin_memory() for k in ("omega_yaw", "velocity_x")} for (dj2, _, _) in keep]
            actrl = [bind_control_host(eng, pj2["x0"], pc["omega_yaw"], pc["velocity_x"], pj2["feet"], it=pj2["iter"], u0=u,
                                       status=o2[1], iters=o2[2], asynchronous=True)
                     for (dj2, pj2, o2), pc, u in zip(keep, pcs, u0s)]
            for j in range(6):
                actrl[j % nh]()
            eng_wait(eng)
            barrier()
            t0 = time.perf_counter()
            for j in range(Ke):
                actrl[j % nh]()
            eng_wait(eng)
            ctrl_async = B_total * Ke / max_over_ranks(time.perf_counter() - t0)
            e2e_ctrl = {"value": max(ctrl_sync, ctrl_async), "unit": "solves/s", "asynchronous_entry_value": ctrl_async,
                        "synchronous_entry_value": ctrl_sync,
                        "h2d_bytes_per_step": B_total * (104 + 16 + 48 + 4), "d2h_bytes_per_step": B_total * (48 + 8), "steps": Ke,
                        "what": "mpc_b200_tron1_control_host: state + (yaw-rate, vx) command + feet + gait clock in, u = U_opt.col(0) "
                                "out (the reference mpcQP constructor's own inputs/outputs); x_ref is generated on the device"}
    else:
        # closed loop end to end: initial states and commands from pinned host memory, final states and counters back
        hp = {k: torch.from_numpy(d[k]).pin_memory() for k in ("x0", "omega_yaw", "velocity_x", "iter")}
        xh = torch.empty((B, 13), dtype=torch.float64).pin_memory(); bh = torch.empty(B, dtype=torch.int32).pin_memory()
        xd = torch.empty((B, 13), dtype=torch.float64, device=dev); oyd = torch.empty(B, dtype=torch.float64, device=dev)
        vxd = torch.empty(B, dtype=torch.float64, device=dev); itd = torch.empty(B, dtype=torch.int32, device=dev)

        def host_rollout():
            xd.copy_(hp["x0"], non_blocking=True); oyd.copy_(hp["omega_yaw"], non_blocking=True)
            vxd.copy_(hp["velocity_x"], non_blocking=True); itd.copy_(hp["iter"], non_blocking=True)
            _, bad, _ = eng.rollout(xd, oyd, vxd, itd, cfg["steps"])
            xh.copy_(xd, non_blocking=True); bh.copy_(bad, non_blocking=True)
            torch.cuda.synchronize()
        host_rollout()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            host_rollout()
        te = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": B_total * cfg["steps"] * K / te, "unit": "solves/s", "h2d_bytes_per_step": B_total * (104 + 8 + 8 + 4),
               "d2h_bytes_per_step": B_total * (104 + 4), "steps": K,
               "path": "pinned host buffers -> cudaMemcpyAsync -> mpc_b200_tron1_rollout_device -> final state and counters back"}
    clocks = sampler.stop() if sampler else None

    # ---- the other two regimes of the headline workload, first-class (config 2 only): short device-timed runs ----
    if args.config == "2" and rank == 0 and not args.no_extras:
        for key, cid in (("standing", "2s"), ("stressed", "2x")):
            c2 = CONFIGS[cid]
            eng2 = Engine(horizon=N, max_batch=B, device=local, Ts=c2["Ts"], mu=c2["mu"])
            pl = []
            for i in range(8):
                dd = synth_batch(c2, B, first=i * B)
                pl.append({k: torch.from_numpy(dd[k]).to(dev) for k in ("x0", "x_ref", "feet", "iter")})
            F2 = torch.empty((B, N, 6), dtype=torch.float64, device=dev)
            s2 = torch.empty(B, dtype=torch.int32, device=dev); i2 = torch.empty(B, dtype=torch.int32, device=dev)
            cl = [eng2.bind_solve(p["x0"], p["x_ref"], p["feet"], it=p["iter"], forces=F2, status=s2, iters=i2,
                                  pipelined=not args.serialized) for p in pl]
            for i in range(8):
                cl[i]()
            eng2.join()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 200
            a0.record()
            for i in range(reps):
                cl[i % 8]()
            eng2.join()
            a1.record()
            torch.cuda.synchronize()
            extra[key] = {"value": B * reps / (a0.elapsed_time(a1) * 1e-3), "unit": "solves/s", "config_id": cid,
                          "mean_iters": float(i2.float().mean().item()), "max_iters": int(i2.max().item()),
                          "unsolved": int((s2 != 0).sum().item()), "workload": workload_name(c2, 1)}
            eng2.close()

    ms_serial_max = max_over_ranks(ms_serial) if ms_serial is not None else None
    max_over_ranks_host = lambda _x: ms_serial_max
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- single-instance latency: host call -> forces on host, B = 1 (horizon 10 configs) ----------------
    latency = None
    if N == 10 and not rollout and not args.no_extras:
        dj, pj, out = keep[0]
        one = {k: torch.from_numpy(dj[k][:1].copy()).pin_memory() for k in ("x0", "x_ref", "feet", "iter")}
        F1 = torch.empty((1, N, 6), dtype=torch.float64).pin_memory()
        s1 = torch.empty(1, dtype=torch.int32).pin_memory(); i1 = torch.empty(1, dtype=torch.int32).pin_memory()

        def lat_of(call, n):
            v = []
            for j in range(n + 200):
                t0 = time.perf_counter()
                call()       # host call -> forces on host
                if j >= 200:
                    v.append(time.perf_counter() - t0)
            return np.array(v if v else [float("nan")]) * 1e6
        walk = lat_of(bind_solve_host(eng, one["x0"], one["x_ref"], one["feet"], it=one["iter"], forces=F1, status=s1, iters=i1),
                      args.latency_calls)
        # BASELINE configs[0]: a single robot STANDING on both feet (gait clock < 0): the double-support class
        one_s = torch.full((1,), -1, dtype=torch.int32).pin_memory()
        stand = lat_of(bind_solve_host(eng, one["x0"], one["x_ref"], one["feet"], it=one_s, forces=F1, status=s1, iters=i1),
                       args.latency_calls)
        latency = {"p50_us": float(np.percentile(stand, 50)), "p99_us": float(np.percentile(stand, 99)), "calls": len(stand),
                   "what": "BASELINE configs[0]: ONE standing robot (both feet in contact), B=1 host call -> forces on host, pinned buffers",
                   "walking_p50_us": float(np.percentile(walk, 50)), "walking_p99_us": float(np.percentile(walk, 99))}

    # ---- roofline of the dominant kernel of the step ---------------------------------------------------
    peaks = load_peaks()
    fp64_peak = measure_fp64_peak(local)
    kernel_ms = ms / K / (cfg["steps"] if rollout else 1)       # per batch-solve (rollout: per control step)
    if cfg["standing"]:
        nc, gthreads, kname = 6 * N, {10: 64, 20: 64, 50: 32}[N], {10: "tron1_solve_kernel<10,60,2,2,3,INDIRECT>", 20: "tron1_solve_kernel<20,120,2,2,1,INDIRECT>", 50: "tron1_solve_kernel<50,300,1,1,8,DIRECT,RICCATI>"}[N]
    else:
        nc, gthreads = 3 * N, {10: 32, 20: 64, 50: 32}[N]
        kname = {10: "tron1_solve_kernel<10,30,1,4,4,DIRECT,persistent>", 20: "tron1_solve_kernel<20,60,2,2,2,DIRECT,persistent>",
                 50: "tron1_solve_kernel<50,300,1,1,8,DIRECT,RICCATI> (one warp per instance, Riccati sweeps)"}[N]
        if rollout:
            kname = "tron1_rollout_kernel<10,30,1,4,4>"
    f_exec, f_tensor = structured_flops(N, nc, max(mean_iters, 1.0), gthreads)
    f_dense = dense_equiv_flops(N, mean_iters)
    achieved_tf = f_exec * B / (kernel_ms * 1e-3) / 1e12
    hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
    achieved_gbs = algorithmic_bytes(N) * B / (kernel_ms * 1e-3) / 1e9
    traffic, ncu_exec = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get("configs", {}).get(args.config)
        if ent and ent.get("batch") == B:
            traffic = ent.get("dram_bytes_per_launch")
            if ent.get("executed_fp64_flops_per_launch"):
                per_launch = B * (cfg["steps"] if rollout else 1)      # solves one launch performs
                ncu_exec = {"flops_per_solve": ent["executed_fp64_flops_per_launch"] / per_launch, "source": ent.get("source"),
                            "counts": "thread-level DFMA x2 + DMUL + DADD (DMMA is not in these counters)",
                            "formula_over_ncu": (f_exec - f_tensor) / (ent["executed_fp64_flops_per_launch"] / per_launch)}
    except Exception:
        pass

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the box's host cores ---------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cp = CpuPath(cfg, cores)
        rsteps = 20 if rollout else None
        n = cp.size_for(12.0, 1 << 20, rsteps)
        t0 = time.perf_counter()
        done = cp.run(rsteps)
        dt = time.perf_counter() - t0
        cpu = {"value": done / dt, "unit": "solves/s", "cores": cores, "kind": "port",
               "sample": f"{n} config-{args.config} instances (seed {cfg['seed']})" + (f" x {rsteps} control steps" if rsteps else "")
                         + f" in {dt:.1f} s, one solve per thread at a time",
               "note": "restated reference CPU path (Eigen/qpOASES unavailable in the image)"}

    line = {
        "metric": metric_name(cfg), "value": value, "unit": "solves/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg, world), "seed": cfg["seed"], "config_id": args.config, "l2": l2_note,
                   "parallelism": f"instance-sharded x{world}, no collective", "instances_per_gpu": B,
                   "mean_iters": mean_iters, "unsolved": n_bad},
        "clocks": clocks,
        "serialized": None if ms_serial is None else {
            "value": B_total * 1e3 / max_over_ranks_host(ms_serial), "unit": "solves/s", "ms_per_step": ms_serial,
            "what": "the same steps through mpc_b200_tron1_solve_device, every batch stream-ordered behind the previous one "
                    "(value uses mpc_b200_tron1_solve_device_pipelined: independent batches overlap on engine-owned streams)"},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved_tf / fp64_peak if fp64_peak else None, "traffic": traffic,
                     "peak_source": "FP64 DFMA peak measured in this run by mpc_b200_measure_fp64_peak "
                                    "(MEASURED_PEAKS.json has no FP64 entry; DMMA.8x8x4 measures 37.1 TFLOP/s, "
                                    "tools/microbench/chol_dmma_bench.cu)",
                     "flops_per_solve": f_exec, "tensor_flops_per_solve": f_tensor, "kernel": kname, "kernel_ms": kernel_ms,
                     "what": "achieved = FP64 operations the kernel EXECUTES per solve (hand-counted structured formula, "
                             "bench.py:structured_flops, DESIGN.md section 4) x instances per step / CUDA-event time per step "
                             "(consecutive launches overlap on the pipelined entry, so this is the sustained rate of the kernel, "
                             "not one launch in isolation; `serialized` has the isolated-launch time)",
                     "ncu_check": ncu_exec,
                     "dense_equiv": {"flops_per_solve": f_dense, "achieved": f_dense * B / (kernel_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                                     "what": "SURVEY 8d DENSE accounting (n^2 p condensing, n^3/3 Cholesky): how fast the PROBLEM is "
                                             "solved relative to a dense evaluation; can exceed the peak, not a hardware fraction"},
                     "replaced_kernel": None if N != 50 else {
                         "kernel": "tron1_solve_kernel<50,150,8,1,1,DIRECT> / <50,300,16,1,1,INDIRECT> (tiled DMMA Cholesky of the condensed Hessian)",
                         "value": 3.26e6, "double_support_value": 0.65e6, "unit": "solves/s", "flops_per_solve": 1.81e6, "frac": 0.19,
                         "source": "profiles/r2_bench_config4.json, profiles/r2_c4_solve_kernel.json (same box class, same workload)",
                         "what": "the O(n^3) class this kernel replaced as the direct class of horizon 50 (-DMPC_RIC_N50=0 restores it): "
                                 "8x the executed arithmetic at 5x the pipe utilisation, 0.55x the throughput (double support: 0.1x)"}},
        "roofline_hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "bytes_per_solve": algorithmic_bytes(N)},
        "cpu_baseline": cpu,
    }
    if e2e_ctrl:
        line["e2e_controller"] = e2e_ctrl
    if latency:
        line["latency"] = latency
    line.update(extra)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--config", default="2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override the config's batch size")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--latency-calls", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the standing / stressed / latency sections (profiling runs)")
    ap.add_argument("--serialized", action="store_true", help="time the stream-ordered device entry instead of the pipelined one")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["B"] = args.batch
    if cfg["N"] != 10 or cfg["kind"] == "rollout":
        # heavier steps: keep the default run within minutes (a horizon-50 batch is ~2.5 ms, a rollout ~0.13 s)
        args.steps = min(args.steps, 200 if cfg["N"] == 20 else 50)
        args.warmup = min(args.warmup, 10)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_gpu(args, cfg)


if __name__ == "__main__":
    main()
