"""Minimal driver for profiling the whole resident step (sample -> fit -> score -> select -> tail) on config 3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

n, h = 100_000, 65_536
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
K, x1, x2, *_ = make_scene(n, 0.4, seed=0)
eng = _native.get_engine(0)
eng.upload_pairs(x1, x2, K)
for r in range(reps):
    eng.sample_device(r, h)
    best, _, _, poses, num, idx, ok, X = eng.two_view(1.5e-6, 10, "rms", "min_error", 50.0, want_mask=False, want_sed=False)
    print("best", best.index, best.err, best.count_extra, "inliers", num, "vote", poses.best, list(poses.counts))
