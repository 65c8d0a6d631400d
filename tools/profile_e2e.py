"""cProfile of the e2e call (two_view_arrays from pinned host arrays, config 3): where the host time goes."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native, two_view  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

n, h = 100_000, 65_536
K, x1, x2, *_ = make_scene(n, 0.4, seed=0)
pa, pb = _native.pinned_empty((n, 2)), _native.pinned_empty((n, 2))
pa[...] = x1
pb[...] = x2
eng = _native.get_engine(0)


def step(seed):
    return two_view.two_view_arrays(K, pa, pb, 1.5e-6, 10, "rms", h, sampler="device", seed=seed, on_degenerate="skip", engine=eng)


for s in range(3):
    step(s)
t0 = time.perf_counter()
for s in range(20):
    step(s)
print("ms per call: %.3f" % ((time.perf_counter() - t0) / 20 * 1e3))
pr = cProfile.Profile()
pr.enable()
for s in range(20):
    step(s)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
