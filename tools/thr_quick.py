"""K2 time at three thresholds for the current library (SFM_B200_LIB selects a variant build)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

n, h = 100_000, 65_536
K, x1, x2, *_ = make_scene(n, 0.4, seed=0)
eng = _native.get_engine(0)
eng.upload_pairs(x1, x2, K)
eng.sample_device(0, h)
eng.fit(want_E=False)
eng.enable_timing(True)
out = []
for thr in (1.5e-6, 1.5e-5, 1.5e-4, 1.5e-3):
    ts = []
    for r in range(4):
        eng.score(thr, 10, "rms", want_arrays=False)
        t, _ = eng.get_timing()
        ts.append(t["score"])
    out.append("%.1e: %.3f ms" % (thr, min(ts[1:])))
print("   ".join(out))
