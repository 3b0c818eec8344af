"""CPU: the N>1 host path (view-pair sharding, max-over-ranks timing) with world_size=2 over gloo."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, ws, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from gps_gaussian_b200 import shard
    r, w = shard.init(backend="gloo")
    assert (r, w) == (rank, ws)
    units = shard.shard_units(7, r, w)
    seeds = shard.unit_seeds(1314, 8, r)
    shard.barrier()
    thr, total, ms = shard.aggregate_throughput(len(units), 100.0 * (r + 1))     # rank 1 is the slow one
    out.put((r, units, seeds, thr, total, ms))
    shard.finalize()


def test_shard_and_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, u0, s0, thr0, tot0, ms0), (r1, u1, s1, thr1, tot1, ms1) = res
    assert sorted(u0 + u1) == list(range(7)) and abs(len(u0) - len(u1)) <= 1     # every unit exactly once, balanced
    assert s0 == list(range(1314, 1322)) and s1 == list(range(1322, 1330))       # disjoint weak-scaling seeds
    assert tot0 == tot1 == 7 and ms0 == ms1 == 200.0                             # MAX over ranks
    assert abs(thr0 - 35.0) < 1e-9 and thr0 == thr1


def test_shard_units_properties():
    sys.path.insert(0, ROOT)
    from gps_gaussian_b200 import shard
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            parts = [shard.shard_units(n, r, w) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
