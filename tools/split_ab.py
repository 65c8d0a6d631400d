"""K2 time (library stage timer) for the shapes where the item split matters: config 2, config 3, its strong-scaling
shards, config 4.  usage: python tools/split_ab.py   (SFM_B200_LIB selects a variant library)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

eng = _native.get_engine(0)
eng.enable_timing(True)
for n, h in ((10_000, 16_384), (100_000, 65_536), (100_000, 32_768), (100_000, 8_192), (5_000, 2_000), (30_000, 30_000)):
    K, x1, x2, *_ = make_scene(n, 0.4, seed=0)
    eng.upload_pairs(x1, x2, K)
    eng.sample_device(0, h)
    eng.fit(want_E=False)
    ts = []
    for r in range(5):
        eng.score(1.5e-6, 10, "rms", want_arrays=False)
        t, _ = eng.get_timing()
        ts.append(t["score"])
    print(f"{n:7d} x {h:6d}: score {min(ts[1:]):8.4f} ms  -> {n * h / min(ts[1:]) * 1e3:.3e} evals/s")
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
    ts.append(t["score"])
print(f"config4 512 pairs: score {min(ts[1:]):8.4f} ms  -> {P * n * h / min(ts[1:]) * 1e3:.3e} evals/s")
