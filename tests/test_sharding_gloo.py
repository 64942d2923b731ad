"""world_size-2 gloo test of the only multi-rank step of the path: rank-strided sharding of the RVE ids
and the final gather of the per-RVE outputs.  The per-RVE "solve" here is the CPU oracle on a
tiny mesh (checker only), so the test also shows that shards reproduce the unsharded result."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ns, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepbnd_b200.sharding import shard_ids, gather_rows
    ids = shard_ids(ns, rank, world)
    out = _fake_outputs(ids)
    gid, res = gather_rows(ids, out, dst=0)
    if rank == 0:
        q.put((gid, res))
    dist.barrier()
    dist.destroy_process_group()


def _fake_outputs(ids):
    """deterministic per-id rows with the snapshot/tangent shapes"""
    t = np.stack([np.arange(9.0).reshape(3, 3) + 100.0 * i for i in ids]) if len(ids) else np.zeros((0, 3, 3))
    s = np.stack([np.full(162, float(i)) for i in ids]) if len(ids) else np.zeros((0, 162))
    return {"tangent": t, "solutions_S": s}


@pytest.mark.parametrize("ns", [7, 8])
def test_rank_strided_shards_gather_to_the_unsharded_result(ns):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + ns
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ns, q)) for r in range(2)]
    for p in procs:
        p.start()
    gid, res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _fake_outputs(np.arange(ns))
    assert np.array_equal(gid, np.arange(ns))
    assert np.array_equal(res["tangent"], ref["tangent"]) and np.array_equal(res["solutions_S"], ref["solutions_S"])


def test_shard_ids_match_reference_striding():
    from deepbnd_b200.sharding import shard_ids, gather_rows
    ns = 11
    parts = [shard_ids(ns, r, 4) for r in range(4)]
    assert [p.tolist() for p in parts] == [list(range(r, ns, 4)) for r in range(4)]     # simulation_snapshots.py:76
    ids, res = gather_rows(parts[1], {"x": parts[1][:, None] * 2.0})                    # no process group: identity
    assert np.array_equal(ids, parts[1]) and np.array_equal(res["x"][:, 0], 2.0 * parts[1])
