"""cProfile of one config-1 estimate through the reference's list API (where does the host time go?)."""
import cProfile
import os
import pstats
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lib.common.feature import Feature  # noqa: E402
from lib.epipolar.epipolar_ransac import estimate_essential_mat_with_ransac  # noqa: E402
from lib.feature_matching.matching import Match  # noqa: E402
from lib.ransac.ransac import ErrorAggregationMethod  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

n, h = 500, 1000
K, x1, x2, *_ = make_scene(n, 0.3, seed=0)
fa = [Feature(x=float(p[0]), y=float(p[1])) for p in x1]
fb = [Feature(x=float(p[0]), y=float(p[1])) for p in x2]
ms = [Match(a_index=i, b_index=i) for i in range(n)]


def step():
    random.seed(5)
    return estimate_essential_mat_with_ransac(K, fa, fb, ms, 1.5e-6, min_num_extra_inliers=10,
                                              error_aggregation_method=ErrorAggregationMethod.RMS, max_iterations=h)


for _ in range(3):
    step()
t0 = time.perf_counter()
for _ in range(20):
    step()
print("ms per estimate: %.3f" % ((time.perf_counter() - t0) / 20 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
