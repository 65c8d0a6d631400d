"""Time the fit stage (CUDA events inside the library): config 3 (64k fits) and config 4 (512 pairs x 2000 fits)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

eng = _native.get_engine(0)
K, x1, x2, *_ = make_scene(100_000, 0.4, seed=0)
eng.upload_pairs(x1, x2, K)
eng.enable_timing(True)
ts = []
for r in range(6):
    eng.sample_device(r, 65536)
    eng.fit(want_E=False)
    t, _ = eng.get_timing()
    ts.append(t["fit"])
print("config3 fit: %.4f ms" % min(ts[1:]))
P, n, h = 512, 2000, 2000
base = make_scene(n, 0.4, seed=0)
pa = np.concatenate([base[1]] * P)
pb = np.concatenate([base[2]] * P)
off = np.arange(P + 1, dtype=np.int64) * n
Ks = np.stack([base[0]] * P)
ts = []
for r in range(5):
    eng.batch_ransac(pa, pb, off, Ks, h, r, 1.5e-6, 10, "rms")
    t, _ = eng.get_timing()
    ts.append((t["fit"], t["score"]))
print("config4 fit: %.4f ms  score %.4f ms" % min(ts[1:]))
import time
ts = []
for r in range(5):
    t0 = time.perf_counter()
    out = eng.batch_two_view(pa, pb, off, Ks, h, r, 1.5e-6, 10, "rms")
    wall = (time.perf_counter() - t0) * 1e3
    t, _ = eng.get_timing()
    ts.append((wall, t))
wall, t = min(ts[1:], key=lambda x: x[0])
print("config4 batch_two_view (512 pairs, pageable host arrays): wall %.3f ms  stages %s  inliers %d" % (
    wall, {k: round(v, 3) for k, v in t.items()}, len(out["inlier_idx"])))
