"""Minimal driver for profiling: upload -> sample -> fit -> score on one workload (no timing)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

n, h = {"config2": (10_000, 16_384), "config3": (100_000, 65_536)}[sys.argv[1] if len(sys.argv) > 1 else "config3"]
variant = sys.argv[2] if len(sys.argv) > 2 else "screen"
hpt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
group = int(sys.argv[5]) if len(sys.argv) > 5 else 0
K, x1, x2, *_ = make_scene(n, 0.4, seed=0)
eng = _native.get_engine(0)
eng.set_score_variant(variant, hpt, group)
eng.upload_pairs(x1, x2, K)
for r in range(reps):
    eng.sample_device(r, h)
    b, _, _ = eng.ransac_essential(float(os.environ.get("SFM_THR", "1.5e-6")), 10, "rms", want_mask=False, want_sed=False)
    print("best", b.index, b.err, b.count_extra, "invalid", b.num_invalid)
