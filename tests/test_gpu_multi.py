"""Hypothesis-sharded estimate over real NCCL ranks (needs >= 2 GPUs on the box; skipped otherwise): the fused
sharded path (device-to-device all-gather of the selection records + merge kernel + tail) must return, on every rank,
exactly what one GPU returns for the union of the hypotheses."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, hyps_per_rank, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    from structure_from_motion_b200 import _native, distributed
    from structure_from_motion_b200.scenes import make_scene

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = _native.Engine(rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    K, x1, x2, *_ = make_scene(20_000, 0.4, seed=2)
    eng.upload_pairs(x1, x2, K)
    res = []
    for seed in (1, 2):
        r = distributed.two_view_sharded(1.5e-6, 10, "rms", hyps_per_rank, seed, engine=eng, rank=rank, world=world)
        res.append(dict(index=r["index"], err=r["err"], count=r["count"], E=r["E"], owner=r["owner"],
                        num=r["num_inliers"], idx=r["inlier_idx"], ok=r["pass_bits"], X=r["points"],
                        best=int(r["poses"].best), counts=list(r["poses"].counts)))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array(res, dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_two_view_sharded_over_nccl_equals_one_gpu(engine, tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from structure_from_motion_b200.scenes import make_scene

    world, H = min(torch.cuda.device_count(), 8), 4096
    mp.spawn(_worker, args=(world, 29571, H, str(tmp_path)), nprocs=world, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npy", allow_pickle=True) for r in range(world)]
    K, x1, x2, *_ = make_scene(20_000, 0.4, seed=2)
    engine.upload_pairs(x1, x2, K)
    for k, seed in enumerate((1, 2)):
        engine.sample_device(seed, world * H)
        best, _, _, poses, num, idx, ok, X = engine.two_view(1.5e-6, 10, "rms", "min_error", 50.0)
        for r in range(world):
            got = ranks[r][k]
            assert got["index"] == best.index and got["err"] == best.err and got["count"] == best.count_extra
            assert got["owner"] == best.index // H
            assert np.array_equal(got["E"].reshape(9), np.array(best.E))
            assert got["best"] == poses.best
            # the selection record carries the winner's model AND its sample row: every rank - owner or not - forces
            # the same 8 points into the inlier set (ransac.py:76) and applies the same index-0 quirk to the vote
            assert got["num"] == num and np.array_equal(got["idx"], idx) and np.array_equal(got["ok"], ok), r
            assert np.array_equal(got["X"], X, equal_nan=True), r
            assert got["counts"] == list(poses.counts), r


def _native_worker(rank, world, unique_id, hyps_per_rank, out_dir):
    """No torch.distributed anywhere: the communicator and the all-gather live behind the C ABI (sfm_nccl_*)."""
    sys.path.insert(0, ROOT)
    from structure_from_motion_b200 import _native
    from structure_from_motion_b200.scenes import make_scene

    eng = _native.Engine(rank)
    K, x1, x2, *_ = make_scene(20_000, 0.4, seed=2)
    eng.upload_pairs(x1, x2, K)
    eng.nccl_init(rank, world, unique_id)
    res = []
    for seed in (1, 2):
        eng.two_view_sharded(seed, hyps_per_rank, 1.5e-6, 10, "rms", "min_error", 50.0)
        best, owner, poses, num, idx, ok, X = eng.sharded_fetch()
        res.append(dict(index=int(best.index), err=float(best.err), count=int(best.count_extra), E=np.array(best.E),
                        owner=owner, num=num, idx=idx, ok=ok, X=X, best=int(poses.best), counts=list(poses.counts)))
    np.save(os.path.join(out_dir, f"native{rank}.npy"), np.array(res, dtype=object), allow_pickle=True)
    eng.nccl_destroy()
    eng.close()


def test_native_nccl_sharded_equals_one_gpu(engine, tmp_path):
    """sfm_nccl_unique_id / sfm_nccl_init / sfm_two_view_sharded / sfm_sharded_fetch: the hypothesis-sharded estimate
    through the C ABI alone, every rank bit-identical to one GPU scoring the union."""
    import multiprocessing as mp

    from structure_from_motion_b200 import _native
    from structure_from_motion_b200.scenes import make_scene

    ndev = _native.load_library().sfm_device_count()
    if ndev < 2:
        pytest.skip("needs two GPUs")
    world, H = min(ndev, 8), 4096
    uid = _native.nccl_unique_id()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_native_worker, args=(r, world, uid, H, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    K, x1, x2, *_ = make_scene(20_000, 0.4, seed=2)
    engine.upload_pairs(x1, x2, K)
    for k, seed in enumerate((1, 2)):
        engine.sample_device(seed, world * H)
        best, _, _, poses, num, idx, ok, X = engine.two_view(1.5e-6, 10, "rms", "min_error", 50.0)
        for r in range(world):
            got = np.load(tmp_path / f"native{r}.npy", allow_pickle=True)[k]
            assert got["index"] == best.index and got["err"] == best.err and got["count"] == best.count_extra, r
            assert got["owner"] == best.index // H and np.array_equal(got["E"], np.array(best.E)), r
            assert got["num"] == num and np.array_equal(got["idx"], idx) and np.array_equal(got["ok"], ok), r
            assert np.array_equal(got["X"], X, equal_nan=True) and got["best"] == poses.best, r
            assert got["counts"] == list(poses.counts), r
