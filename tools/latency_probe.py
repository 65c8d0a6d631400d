"""Fixed per-call cost of the C-ABI entry points on tiny inputs (host wall clock, after warm-up)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

eng = _native.get_engine(0)
if os.environ.get("SFM_PROBE_DEFAULT_STREAM"):
    eng.set_stream(0)
K, x1, x2, *_ = make_scene(200, 0.3, seed=0)


def t(name, fn, reps=200):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    print("%-34s %8.1f us" % (name, (time.perf_counter() - t0) / reps * 1e6))


t("upload_pairs (200)", lambda: eng.upload_pairs(x1, x2, K))
t("sample_device (64)", lambda: eng.sample_device(1, 64))
t("ransac_essential (no mask)", lambda: eng.ransac_essential(1.5e-6, 5, "rms", want_mask=False, want_sed=False))
t("ransac_essential (mask+sed)", lambda: eng.ransac_essential(1.5e-6, 5, "rms"))
t("pose_and_triangulate", lambda: eng.pose_and_triangulate(1.5e-6, 50.0))
off = np.array([0, 200], dtype=np.int64)
t("batch_ransac (1 pair)", lambda: eng.batch_ransac(x1, x2, off, K[None], 64, 1, 1.5e-6, 5, "rms"))
off8 = np.arange(9, dtype=np.int64) * 25
t("batch_ransac (8 pairs of 25)", lambda: eng.batch_ransac(x1, x2, off8, np.stack([K] * 8), 64, 1, 1.5e-6, 2, "rms"))
img = (np.random.default_rng(0).random((64, 64)) * 255).astype(np.uint8)
t("harris_corners 64x64", lambda: eng.harris_corners(img, 20))
f = np.array([[10.0, 10.0], [20.0, 20.0], [30.0, 12.0]])
t("match_brute_force 3x3", lambda: eng.match_brute_force(img, img, f, f, window=5))
t("synchronize", lambda: eng.synchronize())
