"""The list-order array formulation of DistributeOctTree that the CUDA kernel implements (tools/quadtree_proto.py)
against the oracle's linked-list restatement (src/ORBextractor.cc:1006-1287)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
import quadtree_proto as Q  # noqa: E402


def test_array_formulation_matches_list_restatement(oracle):
    rng = np.random.default_rng(1)
    checked = 0
    for trial in range(60):
        W = int(rng.integers(40, 800)); H = int(rng.integers(40, 500))
        if round(W / H) < 1:
            continue
        n = int(rng.integers(1, 900)); N = int(rng.integers(1, 300))
        pts = rng.choice(W * H, size=min(n, W * H), replace=False)
        x = (pts % W).astype(np.int64); y = (pts // W).astype(np.int64)
        sc = rng.integers(7, 40, len(x))
        ref = oracle.distribute_octtree(x.astype(np.float32), y.astype(np.float32), sc.astype(np.float32), 0, W, 0, H, N)
        got = Q.distribute(x, y, sc, np.arange(len(x)), 0, W, 0, H, N)
        assert list(ref) == got, (trial, W, H, n, N)
        checked += 1
    assert checked > 30


def test_quadtree_known_answers(oracle):
    # one key -> itself; N=1 with many keys -> the first pass still runs and returns >= 1 leaves
    r = oracle.distribute_octtree(np.array([5.], np.float32), np.array([7.], np.float32), np.array([9.], np.float32), 0, 100, 0, 60, 10)
    assert list(r) == [0]
    x = np.array([1, 60, 1, 60], np.float32); y = np.array([1, 1, 40, 40], np.float32); s = np.array([1, 2, 3, 4], np.float32)
    r = oracle.distribute_octtree(x, y, s, 0, 100, 0, 60, 1)
    assert sorted(r) == [0, 1, 2, 3]          # nIni = 2 roots, each split once: the result may exceed N and is not trimmed
