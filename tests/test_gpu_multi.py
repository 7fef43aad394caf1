"""GPU tier, >= 2 GPUs (skipped on a single-GPU box): instance sharding runs the CUDA path on every device and gives
bit-identical results to the unsharded solve -- through one engine per device driven from Python (shard.partition, what
bench.py does under torchrun) and through the single-process C-ABI entry mpc_b200_tron1_solve_host_multi."""
import numpy as np
import pytest

from mpc_limx_control_b200 import shard, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def two_gpus():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    return torch


@pytest.mark.parametrize("N,B,standing_every", [(20, 4099, 0), (10, 4096, 7)])
def test_sharded_equals_unsharded(two_gpus, N, B, standing_every):
    """BASELINE configs[2] shape (horizon 20, randomised instances), block-partitioned over cuda:0 / cuda:1"""
    torch = two_gpus
    from mpc_limx_control_b200.engine import Engine, solve_host_multi
    Ts = 0.005
    d = synth.tron1_batch(1002, B, N, Ts)
    if standing_every:
        d["iter"][::standing_every] = -1          # mixed capacity classes on both devices
    ref_eng = Engine(horizon=N, max_batch=B, device=0, Ts=Ts)
    t = {k: torch.from_numpy(np.ascontiguousarray(d[k])).to("cuda:0") for k in ("x0", "x_ref", "feet", "iter")}
    F, st, it = ref_eng.solve(t["x0"], t["x_ref"], t["feet"], it=t["iter"])
    torch.cuda.synchronize(0)
    F, st, it = F.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy()
    assert (st == 0).all()
    G = 2
    engines = [Engine(horizon=N, max_batch=B // G + 1, device=g, Ts=Ts) for g in range(G)]
    # (a) one engine per device, device-resident slices (the torchrun path of bench.py, here in one process)
    outs = []
    for g, e in enumerate(engines):
        s0, cnt = shard.partition(B, G, g)
        dev = f"cuda:{g}"
        with torch.cuda.device(g):
            tg = {k: torch.from_numpy(np.ascontiguousarray(d[k][s0:s0 + cnt])).to(dev) for k in ("x0", "x_ref", "feet", "iter")}
            outs.append(e.solve(tg["x0"], tg["x_ref"], tg["feet"], it=tg["iter"]))
    for g in range(G):
        torch.cuda.synchronize(g)
    Fs = np.concatenate([o[0].cpu().numpy() for o in outs]); ss = np.concatenate([o[1].cpu().numpy() for o in outs])
    assert np.array_equal(Fs, F) and np.array_equal(ss, st)
    assert all(e.launch_count() > 0 for e in engines)            # both devices ran the CUDA path
    # (b) the single-process C-ABI entry with pinned host arrays: every GPU writes its rows of one result array
    pin = {k: torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory() for k in ("x0", "x_ref", "feet", "iter")}
    Fh = torch.empty((B, N, 6), dtype=torch.float64).pin_memory()
    sh = torch.empty(B, dtype=torch.int32).pin_memory(); ih = torch.empty(B, dtype=torch.int32).pin_memory()
    solve_host_multi(engines, pin["x0"], pin["x_ref"], pin["feet"], it=pin["iter"], forces=Fh, status=sh, iters=ih)
    assert np.array_equal(Fh.numpy(), F) and np.array_equal(sh.numpy(), st) and np.array_equal(ih.numpy(), it)
    # pageable arrays take the staged path on every device
    F2, s2, i2 = solve_host_multi(engines, d["x0"], d["x_ref"], d["feet"], it=d["iter"])
    assert np.array_equal(F2, F) and np.array_equal(s2, st)
    for e in engines + [ref_eng]:
        e.close()
