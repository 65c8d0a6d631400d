"""Pair-sharded path carried through the whole hot path: sfm_batch_two_view returns,
per image pair, what the single-pair call returns for that pair (RANSAC E, inlier list, 4-pose cheirality vote,
triangulated inliers) and what the oracle computes from the same winner (lib/epipolar/eight_point.py:65-96, 181-280;
lib/epipolar/triangulation.py:42-62)."""
import numpy as np
import pytest

from oracle import restatement as o
from structure_from_motion_b200.scenes import make_scene

pytestmark = pytest.mark.gpu
THR = 1.5e-6


def _batch(sizes, seed0=100, frac=0.4):
    scenes = [make_scene(max(s, 8), frac, seed=seed0 + p) for p, s in enumerate(sizes)]
    xa = np.concatenate([sc[1][:s] for sc, s in zip(scenes, sizes)])
    xb = np.concatenate([sc[2][:s] for sc, s in zip(scenes, sizes)])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    Ks = np.stack([sc[0] for sc in scenes])
    return scenes, xa, xb, off, Ks


def test_batch_two_view_equals_single_pair_calls(engine):
    sizes = [2000, 1500, 5, 0, 2000, 700, 8, 33, 1025, 1024]
    h, seed, pair0 = 600, 9, 40
    scenes, xa, xb, off, Ks = _batch(sizes)
    out = engine.batch_two_view(xa, xb, off, Ks, h, seed, THR, 10, "rms", pair_id0=pair0)
    assert out["inlier_offsets"][0] == 0 and out["inlier_offsets"][-1] == len(out["inlier_idx"])
    found = 0
    for p, s in enumerate(sizes):
        lo, hi = out["inlier_offsets"][p], out["inlier_offsets"][p + 1]
        if s < 18:  # cannot have 10 extra inliers: no model, nothing downstream
            assert out["best_index"][p] == -1 and lo == hi and out["pose_index"][p] == -2
            assert np.isnan(out["R"][p]).all()
            continue
        a, b = xa[off[p]:off[p + 1]], xb[off[p]:off[p + 1]]
        engine.upload_pairs(a, b, Ks[p])
        engine.sample_device(seed=seed, h=h, stream=pair0 + p)
        best, _, _, poses, num, idx, ok, X = engine.two_view(THR, 10, "rms", "min_error", 50.0)
        assert best.index == out["best_index"][p], p
        if best.index < 0:
            assert lo == hi
            continue
        found += 1
        assert np.array_equal(np.array(best.E).reshape(3, 3), out["E"][p])
        assert hi - lo == num and np.array_equal(out["inlier_idx"][lo:hi], idx), p
        assert np.array_equal(out["pass_bits"][lo:hi], ok), p
        assert int(poses.best) == out["pose_index"][p] and list(poses.counts) == out["counts"][p].tolist(), p
        assert np.array_equal(np.array(poses.R).reshape(4, 3, 3)[poses.best], out["R"][p])
        assert np.array_equal(np.array(poses.t).reshape(4, 3)[poses.best], out["t"][p])
        assert np.array_equal(out["points"][lo:hi], X, equal_nan=True), p
    assert found >= 5


def test_batch_two_view_against_oracle(engine):
    """One pair of a batch against the oracle, from the batch's own winning model: inlier list, pose (up to the
    LAPACK sign ambiguity the candidates share), vote counts and passing set in the reference's list order (samples
    first: the index-0 quirk of eight_point.py:228-230), triangulated points <= 1e-6 relative."""
    sizes = [900, 1200, 640]
    h, seed = 500, 3
    scenes, xa, xb, off, Ks = _batch(sizes, seed0=7, frac=0.3)
    out = engine.batch_two_view(xa, xb, off, Ks, h, seed, THR, 10, "rms", pair_id0=0)
    for p in range(len(sizes)):
        assert out["best_index"][p] >= 0
        K = Ks[p]
        a, b = xa[off[p]:off[p + 1]], xb[off[p]:off[p + 1]]
        nxa, nya = o.k_normalise(a[:, 0], a[:, 1], K)
        nxb, nyb = o.k_normalise(b[:, 0], b[:, 1], K)
        E = out["E"][p]
        # the sample row of the winner: same (seed, pair id) sampler as the single-pair call
        engine.upload_pairs(a, b, K)
        engine.sample_device(seed=seed, h=h, stream=p)
        row = engine.get_table(1, first=int(out["best_index"][p]))[0]
        sed = np.array([o.sed_scalar(nxa[i], nya[i], nxb[i], nyb[i], E) for i in range(len(a))])
        m = sed <= THR
        m[row] = True
        lo, hi = out["inlier_offsets"][p], out["inlier_offsets"][p + 1]
        idx = out["inlier_idx"][lo:hi]
        assert np.array_equal(idx, np.nonzero(m)[0])
        extra = np.nonzero(m)[0]
        lst = np.concatenate([row, extra[~np.isin(extra, row)]])  # ransac.py:76 list order for a table-driven run
        Rr, tr, idx_l, counts_o = o.recover_r_t(nxa[lst], nya[lst], nxb[lst], nyb[lst], E)
        np.testing.assert_allclose(out["R"][p], Rr, atol=1e-9)
        np.testing.assert_allclose(out["t"][p], tr, atol=1e-9)
        assert sorted(out["counts"][p].tolist()) == sorted(int(c) for c in counts_o)
        passing = ((out["pass_bits"][lo:hi] >> out["pose_index"][p]) & 1).astype(bool)
        assert np.array_equal(np.sort(lst[idx_l]), idx[passing])
        pi = idx[passing]
        X_o = o.triangulate_points(a[pi, 0], a[pi, 1], b[pi, 0], b[pi, 1], K, o.tmat(Rr, tr))
        X = out["points"][lo:hi][passing]
        rel = np.linalg.norm(X - X_o, axis=1) / np.linalg.norm(X_o, axis=1)
        assert rel.max() <= 1e-6, rel.max()


def test_pair_pipeline_two_view_equals_single_call(engine):
    from structure_from_motion_b200.distributed import PairPipeline

    sizes = [700, 64, 333, 9, 1200, 500, 8, 410, 77]
    scenes, xa, xb, off, Ks = _batch(sizes, seed0=40, frac=0.35)
    want = engine.batch_two_view(xa, xb, off, Ks, 300, 11, THR, 5, "rms", pair_id0=1000)
    pipe = PairPipeline(depth=2)
    try:
        for chunk in (None, 1, 4, 100):
            got = pipe.batch_two_view(xa, xb, off, Ks, 300, 11, THR, 5, "rms", pair_id0=1000, chunk_pairs=chunk)
            for k in want:
                eq = np.array_equal(got[k], want[k], equal_nan=True) if got[k].dtype.kind == "f" else np.array_equal(got[k], want[k])
                assert eq, (chunk, k)
    finally:
        pipe.close()
