"""One-off fuzz of the pair-sharded batch (sfm_batch_ransac) against the single-pair pipeline: ragged batches with
empty pairs, pairs below eight correspondences, tiny and large hypothesis counts, every selection mode.
usage: python tools/fuzz_batch.py [cases]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(99)
eng = _native.get_engine(0)
bad = 0
for k in range(cases):
    P = int(rng.integers(1, 12))
    sizes = [int(rng.choice([0, 3, 7, 8, 9, 40, 64, 65, 300, 1500, 2500])) for _ in range(P)]
    if sum(sizes) == 0:
        sizes[0] = 50
    h = int(rng.choice([1, 31, 64, 65, 500, 2000]))
    thr = float(10.0 ** rng.uniform(-7, -4))
    min_extra = int(rng.choice([0, 4, 10]))
    agg = ["rms", "sum", "mean", "square"][k % 4]
    sel = ["min_error", "max_inliers", "msac"][k % 3]
    seed, pair0 = int(rng.integers(0, 1000)), int(rng.integers(0, 100))
    scenes = [make_scene(max(s, 8), float(rng.choice([0.0, 0.3, 0.6])), seed=1000 + 13 * k + p) for p, s in enumerate(sizes)]
    xa = np.concatenate([sc[1][:s] for sc, s in zip(scenes, sizes)])
    xb = np.concatenate([sc[2][:s] for sc, s in zip(scenes, sizes)])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    Ks = np.stack([sc[0] for sc in scenes])
    out = eng.batch_ransac(xa, xb, off, Ks, h, seed, thr, min_extra, agg, sel, pair_id0=pair0)
    msg = ""
    for p, s in enumerate(sizes):
        if s < 8:
            if out["best_index"][p] != -1:
                msg = f"pair {p} with {s} correspondences produced a model"
            continue
        a, b = xa[off[p]:off[p + 1]], xb[off[p]:off[p + 1]]
        eng.upload_pairs(a, b, Ks[p])
        eng.sample_device(seed=seed, h=h, stream=pair0 + p)
        best, _, _ = eng.ransac_essential(thr, min_extra, agg, sel, want_mask=False, want_sed=False)
        if best.index != out["best_index"][p]:
            msg = f"pair {p}: winner {out['best_index'][p]} vs {best.index}"
        elif best.index >= 0 and not (best.err == out["best_err"][p] and best.count_extra == out["count_extra"][p]
                                      and np.array_equal(np.array(best.E).reshape(3, 3), out["E"][p])):
            msg = f"pair {p}: winner data differs"
        elif out["num_invalid"][p] != best.num_invalid:
            msg = f"pair {p}: invalid count {out['num_invalid'][p]} vs {best.num_invalid}"
        if msg:
            break
    bad += bool(msg)
    print(f"{k:3d} pairs {P:2d} sizes {sizes} h {h} thr {thr:.1e} {agg} {sel} min_extra {min_extra}  {msg or 'ok'}")
print("mismatches:", bad)
sys.exit(1 if bad else 0)
