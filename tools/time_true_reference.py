"""BASELINE.json configs[0] on the UNMODIFIED reference (build container only: needs /root/reference):
synthetic two-view scene, 500 correspondences, 30 % outliers, 1000 RANSAC iterations through
lib/epipolar/epipolar_ransac.py::estimate_essential_mat_with_ransac, one core, as is.
Writes profiles/r1_true_reference_config1.json.  The reference cannot travel to the GPU box, so this figure is
recorded here; bench.py's reference arm times the (much faster) oracle port on the GPU box's host cores instead."""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_shims  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

ref = reference_shims.load()
F, M = ref.feature.Feature, ref.matching.Match
n, h = 500, 1000
K, x1, x2, *_ = make_scene(n, 0.3, seed=0)
fa = [F(x=float(p[0]), y=float(p[1])) for p in x1]
fb = [F(x=float(p[0]), y=float(p[1])) for p in x2]
ms = [M(a_index=i, b_index=i) for i in range(n)]
random.seed(5)
t0 = time.perf_counter()
e, pairs = ref.epipolar_ransac.estimate_essential_mat_with_ransac(
    K, fa, fb, ms, 1.5e-6, min_num_extra_inliers=10, error_aggregation_method=ref.ransac.ErrorAggregationMethod.RMS,
    max_iterations=h)
dt = time.perf_counter() - t0
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "config1_known_answer.json")))
assert np.array_equal(e, np.array(golden["E"])) and len(pairs) == 23
out = {"what": "unmodified reference, BASELINE.json configs[0] (N=500, 30% outliers, H=1000, thr 1.5e-6, RMS, min_extra 10)",
       "seconds": dt, "evals_per_s": (n - 8) * h / dt, "cores": 1, "python": sys.version.split()[0], "numpy": np.__version__,
       "where": "build container (no GPU)", "inliers": len(pairs), "result_matches_golden": True}
with open(os.path.join(ROOT, "profiles", "r1_true_reference_config1.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
