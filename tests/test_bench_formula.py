"""The roofline numerator bench.py reports (`roofline.flops_per_solve`, a hand-countable structured formula) against what ncu
counted as executed FP64 work for the hot kernel of every config (profiles/traffic.json, written by tools/ncu_summary.py from
the committed `ncu --set full` captures).  CPU tier: no GPU, no oracle."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

# group size (threads per instance) of the hot kernel and the mean active-face solves per instance of each bench workload
HOT = {"2": dict(threads=32, iters=1.003), "2s": dict(threads=64, iters=1.0), "3": dict(threads=64, iters=1.01),
       "4": dict(threads=32, iters=1.0093), "5": dict(threads=32, iters=1.0)}


@pytest.mark.parametrize("cid", ["2", "2s", "3", "4", "5"])
def test_structured_flop_formula_matches_ncu(cid):
    ent = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["configs"][cid]
    cfg = bench.CONFIGS[cid]
    assert ent["batch"] == cfg["B"]
    solves = cfg["B"] * (cfg.get("steps") or 1)
    ncu = ent["executed_fp64_flops_per_launch"] / solves           # thread-level DFMA x 2 + DMUL + DADD (DMMA is not in these counters)
    N = cfg["N"]
    nc = 6 * N if cfg["standing"] else 3 * N
    total, tensor = bench.structured_flops(N, nc, HOT[cid]["iters"], HOT[cid]["threads"])
    assert 0.9 < (total - tensor) / ncu < 1.1, (cid, total, tensor, ncu)


def test_riccati_formula_orders():
    """Horizon 50: the Riccati class executes an order of magnitude less than a dense evaluation of the same problem."""
    trot, _ = bench.structured_flops(50, 150, 1.0, 32)
    dsup, _ = bench.structured_flops(50, 300, 1.0, 32)
    assert 2.0e5 < trot < 2.6e5 and 4.0e5 < dsup < 5.0e5
    assert bench.dense_equiv_flops(50, 1.0) > 100 * dsup
