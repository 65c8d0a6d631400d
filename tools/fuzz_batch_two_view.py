"""Fuzz of sfm_batch_two_view: ragged batches (empty, too short, long pairs) against the single-pair call on the same
(seed, pair id) tables - every output bit-identical.   usage: python tools/fuzz_batch_two_view.py [cases]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(5)
eng = _native.get_engine(0)
bad = 0
for k in range(cases):
    P = int(rng.integers(1, 9))
    sizes = [int(rng.choice([0, 3, 8, 9, 40, 64, 300, 1023, 1024, 1025, 2500, 5000])) for _ in range(P)]
    if sum(sizes) == 0:
        sizes[0] = 50
    h = int(rng.choice([33, 100, 300]))
    thr = float(10.0 ** rng.uniform(-7, -3))
    min_extra = int(rng.choice([0, 5, 10]))
    agg = ["rms", "sum", "mean", "square"][k % 4]
    scenes = [make_scene(max(s, 8), float(rng.choice([0.0, 0.3, 0.5])), seed=900 + 10 * k + p) for p, s in enumerate(sizes)]
    xa = np.concatenate([sc[1][:s] for sc, s in zip(scenes, sizes)])
    xb = np.concatenate([sc[2][:s] for sc, s in zip(scenes, sizes)])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    Ks = np.stack([sc[0] for sc in scenes])
    out = eng.batch_two_view(xa, xb, off, Ks, h, 3 + k, thr, min_extra, agg, pair_id0=17)
    msg = ""
    for p, s in enumerate(sizes):
        lo, hi = out["inlier_offsets"][p], out["inlier_offsets"][p + 1]
        if s < 8:
            if out["best_index"][p] != -1 or lo != hi:
                msg = f"pair {p}: model for {s} correspondences"
            continue
        eng.upload_pairs(xa[off[p]:off[p + 1]], xb[off[p]:off[p + 1]], Ks[p])
        eng.sample_device(3 + k, h, stream=17 + p)
        best, _, _, poses, num, idx, ok, X = eng.two_view(thr, min_extra, agg, "min_error", 50.0)
        if best.index != out["best_index"][p]:
            msg = f"pair {p}: winner {out['best_index'][p]} vs {best.index}"
        elif best.index >= 0 and not (hi - lo == num and np.array_equal(out["inlier_idx"][lo:hi], idx)
                                      and np.array_equal(out["pass_bits"][lo:hi], ok)
                                      and np.array_equal(out["points"][lo:hi], X, equal_nan=True)
                                      and int(poses.best) == out["pose_index"][p]
                                      and list(poses.counts) == out["counts"][p].tolist()
                                      and best.err == out["best_err"][p]):
            msg = f"pair {p}: tail differs"
    bad += bool(msg)
    print(f"{k:3d} sizes {sizes} h {h} thr {thr:8.2e} {agg} min_extra {min_extra}: {msg or 'ok'}")
print("mismatches:", bad)
sys.exit(1 if bad else 0)
