#!/usr/bin/env python
"""bench.py -- batched MPC solves/sec (TRON1, horizon 10) on N B200s + p50 single-solve latency.

One "step" = one pass of the hot path (linearise -> discretise -> condense -> QP solve -> forces)
over one batch of B synthetic instances per GPU (BASELINE.json configs[1]: B=4096, N=10, trot
contact schedule from the gait clock, friction pyramid).  Instances are independent, so N GPUs
each run their own B instances (weak scaling, no collective on the solve path; torch.distributed
is used only for the barrier and the max-over-ranks time).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

--impl reference times the reference's CPU path (the oracle port of QPSolver/mpcQP + a cold-start
active-set QP; Eigen/qpOASES are not installable here) on all host cores, same workload."""
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


SEED, HORIZON, TS = 1001, 10, 0.005
METRIC = "batched MPC solves/sec (TRON1, N=10)"


def algorithmic_flops(N, iters):
    """SURVEY.md 8d fixed accounting (n = 6N, p = 13(N+1))."""
    n, p = 6 * N, 13 * (N + 1)
    return 6422 * N + (n * n * p + n * p) + (26 * p + 2 * p + 2 * p * n) + n ** 3 / 3 + iters * (4 * n * n + 10 * n + 32 * N)


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


def cpu_reference(nthreads, target_seconds, B_cap):
    """Times the oracle port (restated reference CPU path) on `nthreads` host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from mpc_limx_control_b200 import synth
    p = O.tron1_defaults(Ts=TS)
    probe = max(nthreads * 4, 32)
    d = synth.tron1_batch(SEED, probe, HORIZON, TS)
    c = np.stack([O.contact_schedule(int(i), HORIZON) for i in d["iter"]])
    O.tron1_solve_batch(p, HORIZON, d["x0"], d["x_ref"], d["feet"], c, nthreads)   # warm-up
    t = time.perf_counter()
    O.tron1_solve_batch(p, HORIZON, d["x0"], d["x_ref"], d["feet"], c, nthreads)
    rate = probe / (time.perf_counter() - t)
    n = int(min(B_cap, max(probe, rate * target_seconds)))
    d = synth.tron1_batch(SEED, n, HORIZON, TS)
    c = np.stack([O.contact_schedule(int(i), HORIZON) for i in d["iter"]])
    return O, p, d, c, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample per step so that the whole K+W run ends within a few minutes
    O, p, d, c, n = cpu_reference(cores, min(2.0, 150.0 / (args.steps + args.warmup)), args.batch)
    for _ in range(args.warmup):
        O.tron1_solve_batch(p, HORIZON, d["x0"], d["x_ref"], d["feet"], c, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        F, st, it = O.tron1_solve_batch(p, HORIZON, d["x0"], d["x_ref"], d["feet"], c, cores)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = f"{n} of the {args.batch} config-2 instances per step (seed {SEED}), {cores} threads, one solve per thread at a time"
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"TRON1 convex MPC, horizon {HORIZON}, Ts {TS}, trot contact schedule + friction pyramid, "
                               f"B={args.batch} per step (bounded sample: {n})", "seed": SEED},
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "restated reference CPU path (Eigen/qpOASES unavailable): dense condensing + cold-start active set"},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from mpc_limx_control_b200 import synth
    from mpc_limx_control_b200.engine import Engine, measure_fp64_peak

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
    N, B, K, W = HORIZON, args.batch, args.steps, args.warmup

    # rotating pool of distinct input batches, larger than the 126 MB L2, resident in HBM
    per_batch_in = B * (104 + 104 * (N + 1) + 48 + 4)
    pool_n = max(2, int(np.ceil(160e6 / per_batch_in)))
    pool = []
    for i in range(pool_n):
        d = synth.tron1_batch(SEED, B, N, TS, first=(rank * pool_n + i) * B)
        pool.append({k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "x_ref", "feet", "iter")})
    eng = Engine(horizon=N, max_batch=B, device=local, Ts=TS)
    forces = torch.empty((B, N, 6), dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)

    # one pre-bound C-ABI call per pool entry: the timed loop issues mpc_b200_tron1_solve_device and nothing else
    calls = [eng.bind_solve(p["x0"], p["x_ref"], p["feet"], it=p["iter"], forces=forces, status=status, iters=iters) for p in pool]

    def step(i):
        calls[i % pool_n]()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    n_bad = int((status != 0).sum().item())
    mean_iters = float(iters.float().mean().item())
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside) -------
    d = synth.tron1_batch(SEED, B, N, TS, first=rank * B)
    pin = {k: torch.from_numpy(d[k]).pin_memory() for k in ("x0", "x_ref", "feet", "iter")}
    Fh = torch.empty((B, N, 6), dtype=torch.float64).pin_memory()
    sh = torch.empty(B, dtype=torch.int32).pin_memory()
    ih = torch.empty(B, dtype=torch.int32).pin_memory()
    Ke = min(500, max(10, K // 4))
    from mpc_limx_control_b200.engine import bind_solve_host, bind_control_host
    host_call = bind_solve_host(eng, pin["x0"], pin["x_ref"], pin["feet"], it=pin["iter"], forces=Fh, status=sh, iters=ih)
    # a rotating set of distinct pinned input/output batches: no step can be served from a copy of the previous step's
    # bytes in any cache on either side of PCIe
    host_calls = [host_call]
    for j in range(1, 4):
        dj = synth.tron1_batch(SEED, B, N, TS, first=(world * (j + 1) + rank) * B)
        pj = {k: torch.from_numpy(dj[k]).pin_memory() for k in ("x0", "x_ref", "feet", "iter")}
        host_calls.append(bind_solve_host(eng, pj["x0"], pj["x_ref"], pj["feet"], it=pj["iter"],
                                          forces=torch.empty((B, N, 6), dtype=torch.float64).pin_memory(),
                                          status=torch.empty(B, dtype=torch.int32).pin_memory(),
                                          iters=torch.empty(B, dtype=torch.int32).pin_memory()))
    for j in range(max(4, W // 4)):
        host_calls[j % 4]()
    barrier()
    t0 = time.perf_counter()
    for j in range(Ke):
        host_calls[j % 4]()      # one mpc_b200_tron1_solve_host: inputs over PCIe, solve, results back, sync (pinned host buffers)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * Ke / float(te.item())
    e2e_path = "zero-copy (kernel reads/writes the pinned host buffers over PCIe)" if eng.last_host_path() else "staged copies"
    # the same call forced onto the staged-copy path (what a caller with pageable buffers gets), for comparison
    eng.set_host_mode(Engine.HOST_STAGED)
    for _ in range(3):
        host_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(5, Ke // 4)):
        host_call()
    torch.cuda.synchronize()
    ts_ = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
    e2e_staged_value = world * B * max(5, Ke // 4) / float(ts_.item())
    eng.set_host_mode(Engine.HOST_AUTO)
    # controller-shaped host call (command in, first-step force out: the reference mpcQP's own I/O)
    pin_c = {k: torch.from_numpy(d[k]).pin_memory() for k in ("omega_yaw", "velocity_x")}
    u0h = torch.empty((B, 6), dtype=torch.float64).pin_memory()
    ctrl_call = bind_control_host(eng, pin["x0"], pin_c["omega_yaw"], pin_c["velocity_x"], pin["feet"], it=pin["iter"], u0=u0h,
                                  status=sh, iters=ih)
    for _ in range(max(3, W // 4)):
        ctrl_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        ctrl_call()
    torch.cuda.synchronize()
    tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    e2e_ctrl_value = world * B * Ke / float(tc.item())
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- single-instance latency: host call -> forces on host, B = 1 --------------------------------
    lat = []
    one = {k: torch.from_numpy(d[k][:1].copy()).pin_memory() for k in ("x0", "x_ref", "feet", "iter")}
    F1 = torch.empty((1, N, 6), dtype=torch.float64).pin_memory()
    s1 = torch.empty(1, dtype=torch.int32).pin_memory(); i1 = torch.empty(1, dtype=torch.int32).pin_memory()
    one_call = bind_solve_host(eng, one["x0"], one["x_ref"], one["feet"], it=one["iter"], forces=F1, status=s1, iters=i1)
    for j in range(args.latency_calls + 200):
        t0 = time.perf_counter()
        one_call()       # host call -> forces on host
        if j >= 200:
            lat.append(time.perf_counter() - t0)
    lat = np.array(lat if lat else [float("nan")]) * 1e6
    # BASELINE configs[0]-style single robot STANDING on both feet (gait clock < 0): the double-support class
    lat_s = []
    one_s = torch.full((1,), -1, dtype=torch.int32).pin_memory()
    stand_call = bind_solve_host(eng, one["x0"], one["x_ref"], one["feet"], it=one_s, forces=F1, status=s1, iters=i1)
    for j in range(args.latency_calls // 4 + 200):
        t0 = time.perf_counter()
        stand_call()
        if j >= 200:
            lat_s.append(time.perf_counter() - t0)
    lat_s = np.array(lat_s if lat_s else [float("nan")]) * 1e6

    # ---- roofline of the dominant (only) kernel of the step --------------------------------------------
    peaks = load_peaks()
    fp64_peak = measure_fp64_peak(local)
    kernel_ms = ms / K                      # one kernel launch per step, timed with CUDA events on its stream
    flops = algorithmic_flops(N, mean_iters) * B
    achieved_tf = flops / (kernel_ms * 1e-3) / 1e12
    hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
    achieved_gbs = algorithmic_bytes(N) * B / (kernel_ms * 1e-3) / 1e9
    traffic, executed = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get("dram_bytes_per_launch")
        ex = tj.get("executed_fp64") or {}
        if "flops_per_launch" in ex and B == tj.get("batch"):
            # executed FP64 work from the committed ncu capture (thread-level DFMA x2 + DMUL + DADD), same batch
            executed = {"flops_per_solve": ex["flops_per_launch"] / B,
                        "achieved": ex["flops_per_launch"] / (kernel_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                        "frac": ex["flops_per_launch"] / (kernel_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                        "source": tj.get("source")}
    except Exception:
        pass

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the box's host cores ---------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        O, p, dd, cc, n = cpu_reference(cores, 12.0, 1 << 20)
        t0 = time.perf_counter()
        O.tron1_solve_batch(p, N, dd["x0"], dd["x_ref"], dd["feet"], cc, cores)
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": "solves/s", "cores": cores, "kind": "port",
               "sample": f"{n} config-2 instances (seed {SEED}) in {dt:.1f} s, one solve per thread at a time",
               "note": "restated reference CPU path (Eigen/qpOASES unavailable in the image)"}

    line = {
        "metric": METRIC, "value": world * B * K / (ms_max * 1e-3), "unit": "solves/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"TRON1 convex MPC, horizon {N}, Ts {TS}, B={B} instances per GPU per step, trot contact "
                               "schedule from the gait clock, friction pyramid mu=0.5, f_max=2mg (BASELINE configs[1])",
                   "seed": SEED, "l2": f"rotating pool of {pool_n} distinct input batches ({pool_n * per_batch_in / 1e6:.0f} MB) > 126 MB L2",
                   "parallelism": f"instance-sharded x{world}, no collective", "mean_iters": mean_iters, "unsolved": n_bad},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": B * (104 + 104 * (N + 1) + 48 + 4),
                "d2h_bytes_per_step": B * (48 * N + 8), "steps": Ke, "path": e2e_path,
                "host_batches": "4 distinct pinned input/output batches in rotation",
                "staged_copies_value": e2e_staged_value},
        "e2e_controller": {"value": e2e_ctrl_value, "unit": "solves/s", "h2d_bytes_per_step": B * (104 + 16 + 48 + 4),
                           "d2h_bytes_per_step": B * (48 + 8), "steps": Ke,
                           "what": "mpc_b200_tron1_control_host: state + (yaw-rate, vx) command + feet + gait clock in, "
                                   "u = U_opt.col(0) out (the reference mpcQP constructor's own inputs/outputs); x_ref is "
                                   "generated on the device"},
        "gpu_launches": int(launches),
        "latency": {"p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)), "calls": len(lat),
                    "what": "B=1 host call -> forces on host (pinned buffers)",
                    "standing_p50_us": float(np.percentile(lat_s, 50)), "standing_p99_us": float(np.percentile(lat_s, 99))},
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved_tf / fp64_peak if fp64_peak else None, "traffic": traffic,
                     "peak_source": "FP64 DFMA peak measured in this run by mpc_b200_measure_fp64_peak "
                                    "(MEASURED_PEAKS.json has no FP64 entry)",
                     "flops_per_solve": algorithmic_flops(N, mean_iters), "kernel": "tron1_solve_kernel<10,30,1,4,4,false>",
                     "kernel_ms": kernel_ms,
                     "note": "achieved uses SURVEY 8d's DENSE accounting (n^2 p condensing, n^3/3 Cholesky); the kernel's structured "
                             "condensing executes ~10x fewer FLOPs, so frac can exceed 1 -- `executed` is the pipe-level view",
                     "executed": executed},
        "roofline_hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "bytes_per_solve": algorithmic_bytes(N)},
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--latency-calls", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
