"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: contiguous database shards, local top-2 with
global indices, all-gather, (distance, index) merge == unsharded result; frame blocks cover the batch exactly once."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from oracle import oracle as O
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    T, Q = pkg.synth.knn_database(3001, 257)
    T[2999] = T[10]; T[1500] = T[10]                       # ties that straddle shard boundaries
    b = pkg.sharding.shard_bounds(len(T), world)

    def local(q, t, base):
        i, d = O.knn2(q, t)
        i = np.where(i >= 0, i + base, -1).astype(np.int32)
        return i, d

    def gather(x):
        parts = [torch.empty_like(torch.from_numpy(x)) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(x)))
        return [p.numpy() for p in parts]

    idx, dd = pkg.sharding.sharded_knn2(local, pkg.sharding.merge_top2_numpy, gather, Q, T[b[rank]:b[rank + 1]], b[rank], world)
    ref_i, ref_d = O.knn2(Q, T)
    ok = np.array_equal(idx, ref_i) and np.array_equal(dd, ref_d)
    f0, f1 = pkg.sharding.frames_for_rank(1000, rank, world)
    cover = torch.zeros(1000, dtype=torch.int32); cover[f0:f1] = 1
    dist.all_reduce(cover)
    ok = ok and bool((cover == 1).all())
    dist.barrier()
    dist.destroy_process_group()
    out[rank] = int(ok)


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_knn_merge_matches_unsharded(world):
    ctx = mp.get_context('spawn')
    out = ctx.Array('i', [0] * world)
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * world


def test_shard_bounds_and_merge_known_answers(pkg):
    sh = pkg.sharding
    assert sh.shard_bounds(10, 3) == [0, 3, 6, 10]
    assert sh.shard_bounds(8, 8) == list(range(9))
    i, d = sh.merge_top2_numpy([np.array([[5, 9]], np.int32), np.array([[2, -1]], np.int32)],
                               [np.array([[10, 12]], np.int32), np.array([[10, 257]], np.int32)])
    assert i.tolist() == [[2, 5]] and d.tolist() == [[10, 10]]      # equal distance -> lower global index first
