"""CPU tier: the N>1 path (instance sharding + final gather) on a world_size-2 gloo group.
Each rank solves its contiguous slice with the host emulation of the device source (the GPU is
not needed to check the partition / gather logic); rank 0 compares with the unsharded solve."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from mpc_limx_control_b200 import shard


def test_partition_covers_batch():
    for B in (1, 2, 7, 64, 4096, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.partition(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == B
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, N, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import emul_lib as E
    import oracle_lib as O
    from mpc_limx_control_b200 import synth, shard as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = sh.partition(B, world, rank)
    d = synth.tron1_batch(1001, count, N, 0.005, first=start)      # counter-based: slice == slice of the full batch
    p = E.default_params()
    F = np.zeros((count, N, 6)); st = np.zeros((count,), np.int32)
    for b in range(count):
        c = O.contact_schedule(int(d["iter"][b]), N)
        F[b], st[b], _ = E.solve(p, N, d["x0"][b], d["x_ref"][b], d["feet"][b], c)
    Fall = sh.gather_rows(F, B)
    sall = sh.gather_rows(st, B)
    if rank == 0:
        np.savez(out_path, F=Fall, st=sall)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_solve(tmp_path):
    import emul_lib as E
    import oracle_lib as O
    from mpc_limx_control_b200 import synth
    B, N, world = 13, 10, 2
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), B, N, out), nprocs=world, join=True)
    got = np.load(out)
    d = synth.tron1_batch(1001, B, N, 0.005)
    p = E.default_params()
    ref = np.stack([E.solve(p, N, d["x0"][b], d["x_ref"][b], d["feet"][b], O.contact_schedule(int(d["iter"][b]), N))[0]
                    for b in range(B)])
    assert got["F"].shape == (B, N, 6) and np.array_equal(got["F"], ref)
    assert (got["st"] == 0).all()
