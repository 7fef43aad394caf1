"""Secondary measurements on one GPU for the BASELINE configs that are not the headline bench line.
Device-resident inputs, CUDA events, >=3 warm-ups.  Prints one JSON object per config.
usage (GPU box): python tools/bench_configs.py [--quick]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mpc_limx_control_b200 import synth
from mpc_limx_control_b200.engine import Engine

quick = "--quick" in sys.argv
dev = torch.device("cuda", 0)


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def solve_config(name, B, N, seed, Ts=0.005, standing=False, scale=1.0, mu=0.5, reps=20):
    pool = []
    per = B * (104 + 104 * (N + 1) + 48 + 4)
    for i in range(max(2, int(np.ceil(160e6 / per)))):
        d = synth.tron1_batch(seed, B, N, Ts, first=i * B, standing=standing)
        d["x0"][:, [0, 1, 6, 7, 8, 9, 10, 11]] *= scale
        pool.append({k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "x_ref", "feet", "iter")})
    eng = Engine(horizon=N, max_batch=B, Ts=Ts, mu=mu)
    F = torch.empty((B, N, 6), dtype=torch.float64, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
    k = [0]

    def step():
        p = pool[k[0] % len(pool)]; k[0] += 1
        eng.solve(p["x0"], p["x_ref"], p["feet"], it=p["iter"], forces=F, status=st, iters=it)
    ms = timed(step, reps)
    out = dict(config=name, B=B, N=N, Ts=Ts, standing=standing, state_scale=scale, mu=mu, ms_per_batch=ms,
               solves_per_s=B / (ms * 1e-3), mean_iters=float(it.float().mean()), max_iters=int(it.max()),
               uncertified=int((st != 0).sum()))
    eng.close()
    print(json.dumps(out)); sys.stdout.flush()
    return out


def rollout_config(name, B, N, steps, seed, Ts=0.005):
    d = synth.tron1_batch(seed, B, N, Ts)
    eng = Engine(horizon=N, max_batch=B, Ts=Ts)
    x0 = torch.from_numpy(d["x0"]).to(dev)
    oy = torch.from_numpy(d["omega_yaw"]).to(dev); vx = torch.from_numpy(d["velocity_x"]).to(dev); it0 = torch.from_numpy(d["iter"]).to(dev)
    res = {}

    def run():
        x = x0.clone()
        _, bad, its = eng.rollout(x, oy, vx, it0, steps)
        res["bad"], res["its"], res["x"] = bad, its, x
    ms = timed(run, 2, warm=1)
    out = dict(config=name, B=B, N=N, steps=steps, ms_per_rollout=ms, solves_per_s=B * steps / (ms * 1e-3),
               us_per_control_step=ms * 1e3 / steps, uncertified_steps=int(res["bad"].sum()),
               mean_iters_per_step=float(res["its"].float().mean()) / steps, finite=bool(torch.isfinite(res["x"]).all()))
    eng.close()
    print(json.dumps(out)); sys.stdout.flush()
    return out


if __name__ == "__main__":
    solve_config("2: B=4096 N=10 trot (headline)", 4096, 10, 1001)
    solve_config("2s: B=4096 N=10 standing (double support, large capacity class)", 4096, 10, 1001, standing=True)
    solve_config("2x: B=4096 N=10 stress (state x5, mu=0.3, Ts=0.02)", 4096, 10, 1001, Ts=0.02, scale=5.0, mu=0.3)
    solve_config("3: B=65536 N=20 randomised (one GPU's worth of the 8-GPU shard = 8192)", 8192, 20, 1002)
    if not quick:
        solve_config("3f: B=65536 N=20 on one GPU", 65536, 20, 1002, reps=5)
    solve_config("4: B=8192 N=50 long horizon, trot (n = 150, factor in shared memory)", 8192, 50, 1003, reps=3)
    if not quick:
        solve_config("4s: B=1024 N=50 long horizon, double support (n = 300, factor in global memory)", 1024, 50, 1003, standing=True, reps=2)
    rollout_config("5: closed loop, 2048 instances (16384/8 GPUs) x 1000 steps", 2048, 10, 100 if quick else 1000, 1004)
    if not quick:
        rollout_config("5f: closed loop, 16384 instances x 1000 steps on one GPU", 16384, 10, 1000, 1004)
