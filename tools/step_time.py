"""Wall time of the resident fused step (device sampler -> fit -> score -> select -> tail, one blocking call per estimate)
for config 2 and config 3, no L2 flush.  usage: python tools/step_time.py   (SFM_B200_LIB selects a variant library)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

eng = _native.get_engine(0)
for name, n, h, reps in (("config2", 10_000, 16_384, 400), ("config3", 100_000, 65_536, 60), ("2k x 2k", 2_000, 2_000, 400)):
    K, x1, x2, *_ = make_scene(n, 0.4, seed=0)
    eng.upload_pairs(x1, x2, K)
    best = None
    for r in range(3):
        ts = []
        for s in range(reps):
            t0 = time.perf_counter()
            eng.sample_device(s, h)
            out = eng.two_view(1.5e-6, 10, "rms", "min_error", 50.0, want_mask=False, want_sed=False)
            ts.append(time.perf_counter() - t0)
        ts.sort()
        med = ts[len(ts) // 2]
        best = med if best is None else min(best, med)
    print(f"{name:8s}: median step {best * 1e3:8.4f} ms   (winner {out[0].index}, inliers {out[4]})")
