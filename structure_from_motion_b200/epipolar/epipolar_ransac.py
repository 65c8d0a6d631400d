"""Drop-in mirror of lib/epipolar/epipolar_ransac.py (:18-70) — the entry point apps/sfm.py:110 calls."""
from __future__ import annotations

import copy
from functools import partial  # noqa: F401  (kept: callers build the same partials as the reference)
from typing import Tuple

import numpy as np

from .. import two_view
from ..common.feature import Feature
from ..feature_matching.matching import Match
from ..ransac.ransac import ErrorAggregationMethod, fit_with_ransac  # noqa: F401
from .eight_point import estimate_essential_mat, to_normalized_image_coords
from .sed import calculate_symmetric_epipolar_distance

FeaturePair = Tuple[Feature, Feature]


def calculate_sed_inlier_score(e, matching_features: FeaturePair, camera_matrix) -> float:
    """Symmetric epipolar distance of one pixel-coordinate pair under E (epipolar_ransac.py:18-25): both features are
    K-normalised first.  ``fit_with_ransac`` recognises ``partial(calculate_sed_inlier_score, camera_matrix=K)``."""
    normalised = [to_normalized_image_coords(f, camera_matrix) for f in matching_features[:2]]
    return calculate_symmetric_epipolar_distance(feature_a=normalised[0], feature_b=normalised[1], e=e)


def eight_point_model_fitter(matching_features: list, camera_matrix):
    """E from exactly eight pixel-coordinate pairs (epipolar_ransac.py:28-42); ValueError for any other count."""
    count = len(matching_features)
    if count != 8:
        raise ValueError("Eight feature pairs are expected.")
    first, second = zip(*matching_features)
    identity_matches = [Match(a_index=k, b_index=k) for k in range(count)]
    return estimate_essential_mat(camera_matrix=camera_matrix, features_a=list(first), features_b=list(second),
                                  matches=identity_matches)


def estimate_essential_mat_with_ransac(
    camera_matrix,
    features_a: list,
    features_b: list,
    matches: list,
    sed_inlier_threshold: float,
    min_num_extra_inliers: int | None = None,
    error_aggregation_method: ErrorAggregationMethod | None = None,
    max_iterations: int | None = None,
):
    """epipolar_ransac.py:45-70 — returns (E, list of inlier (Feature, Feature) pairs).

    Same sampling (process-global ``random`` state), candidate rule, min-error selection and
    exceptions as the reference; the returned pairs are copies in the reference's order
    (the 8 sample pairs, then the other inliers in that iteration's permutation order).
    """
    pts_a = np.array([[features_a[m.a_index].x, features_a[m.a_index].y] for m in matches],
                     dtype=np.float64).reshape(-1, 2)
    pts_b = np.array([[features_b[m.b_index].x, features_b[m.b_index].y] for m in matches],
                     dtype=np.float64).reshape(-1, 2)
    res = two_view.ransac_essential_arrays(
        camera_matrix, pts_a, pts_b, sed_inlier_threshold, min_num_extra_inliers,
        error_aggregation_method, max_iterations, sampler="reference")
    pairs = []
    for i in res.inlier_indices:
        m = matches[int(i)]
        pairs.append((copy.deepcopy(features_a[m.a_index]), copy.deepcopy(features_b[m.b_index])))
    return res.E, pairs
