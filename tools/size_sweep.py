"""K2 throughput over problem sizes (single pair): looks for partitioning cliffs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

eng = _native.get_engine(0)
eng.enable_timing(True)
print("      N        H     ms     evals/s")
for n, h in [(500, 1000), (500, 16384), (2000, 2000), (2000, 65536), (10000, 2000), (10000, 16384), (10000, 262144),
             (50000, 16384), (100000, 4096), (100000, 65536), (300000, 65536), (1048576, 16384)]:
    K, x1, x2, *_ = make_scene(n, 0.4, seed=0)
    eng.upload_pairs(x1, x2, K)
    eng.sample_device(0, h)
    eng.fit(want_E=False)
    ts = []
    for r in range(4):
        eng.score(1.5e-6, 10, "rms", want_arrays=False)
        t, _ = eng.get_timing()
        ts.append(t["score"])
    print(f"{n:8d} {h:8d} {min(ts[1:]):7.3f}  {n * h / min(ts[1:]) * 1e3:.3e}")
