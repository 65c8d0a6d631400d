"""Where the one-sided and the two-sided fp64 screens cross (config 3, K2 alone), against the pass rate of the one-sided
test that the AUTO pilot measures (kPilotFullAbove in csrc/sfm_score.cuh)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

n, h = 100_000, 65_536
K, x1, x2, *_ = make_scene(n, 0.4, seed=0)
eng = _native.get_engine(0)
eng.upload_pairs(x1, x2, K)
eng.sample_device(0, h)
eng.fit(want_E=False)
E, valid = eng.get_models()
Kinv = np.linalg.inv(K)
a = (np.c_[x1, np.ones(n)] @ Kinv.T)[:2048]
b = (np.c_[x2, np.ones(n)] @ Kinv.T)[:2048]
Es = E[:: h // 64][:64]
lb = np.einsum("hij,ni->hnj", Es, b)          # E^T b
r = np.einsum("hnj,nj->hn", lb, a)
nb = lb[..., 0] ** 2 + lb[..., 1] ** 2
eng.enable_timing(True)
print("thr       one-sided pass rate   screen ms   full ms")
for thr in (1.5e-5, 3e-5, 5e-5, 7e-5, 1e-4, 1.5e-4):
    rate = float(np.mean(r * r <= thr * nb))
    out = []
    for v in ("screen", "full"):
        eng.set_score_variant(v, 2, 16)
        ts = []
        for _ in range(3):
            eng.score(thr, 10, "rms", want_arrays=False)
            t, _ = eng.get_timing()
            ts.append(t["score"])
        out.append(min(ts[1:]))
    print(f"{thr:8.1e}   {rate:8.4f}            {out[0]:8.3f}   {out[1]:8.3f}")
