"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the hypothesis-sharded merge
(one all-gather, reference selection rule) and the pair/hypothesis partitioning."""
import math
import os
import socket
import sys

import numpy as np
import pytest

from structure_from_motion_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_pairs_rebases_offsets():
    off = np.array([0, 5, 5, 12, 20, 21])
    p0, p1, local = D.shard_pairs(off, 1, 2)
    assert (p0, p1) == (3, 5) and local.tolist() == [0, 8, 9]
    p0, p1, local = D.shard_pairs(off, 0, 2)
    assert (p0, p1) == (0, 3) and local.tolist() == [0, 5, 5, 12]


def test_merge_best_rules():
    E = np.arange(9.0)
    rows = np.stack([
        D.pack_local_best(3e-7, 70000, 25, E),
        D.pack_local_best(2e-7, 140000, 12, E + 1),
        D.pack_local_best(2e-7, 5000, 11, E + 2),   # same error, earlier iteration -> wins (ransac.py:83)
        D.pack_local_best(math.inf, -1, -1, E),      # rank without any candidate
    ])
    owner, err, idx, cnt, Ew = D.merge_best(rows)
    assert (owner, err, idx, cnt) == (2, 2e-7, 5000, 11) and np.array_equal(Ew, (E + 2).reshape(3, 3))
    owner, err, idx, cnt, _ = D.merge_best(rows, "max_inliers")
    assert (owner, idx, cnt) == (0, 70000, 25)
    none = np.stack([D.pack_local_best(math.inf, -1, -1, E)] * 3)
    assert D.merge_best(none)[0] == -1
    assert D.all_gather_best(rows[0]).shape == (1, 12)  # no process group: identity


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from structure_from_motion_b200 import distributed as DD

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        H = 1000
        # a synthetic per-hypothesis error table shared by construction: each rank "scores" its shard
        rng = np.random.default_rng(7)
        err = rng.random(world * H)
        err[1234 % (world * H)] = err.min()  # an exact tie across ranks: the lower global index must win
        lo, hi = DD.shard_range(world * H, rank, world)
        j = int(np.argmin(err[lo:hi]))
        E = np.full(9, float(rank))
        rows = DD.all_gather_best(DD.pack_local_best(err[lo + j], lo + j, 20 + rank, E))
        owner, e, idx, cnt, Ew = DD.merge_best(rows)
        exp = int(np.flatnonzero(err == err.min())[0])
        q.put((rank, owner, idx, exp, float(e) == float(err.min()), float(Ew[0, 0])))
    finally:
        dist.destroy_process_group()


def test_hypothesis_sharded_merge_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, owner, idx, exp, err_ok, e00 in out:
        assert idx == exp and err_ok           # every rank agrees on the global winner
        assert owner == (0 if exp < 1000 else 1) and e00 == float(owner)
