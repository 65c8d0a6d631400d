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
    """epipolar_ransac.py:18-25."""
    feature_a = to_normalized_image_coords(matching_features[0], camera_matrix)
    feature_b = to_normalized_image_coords(matching_features[1], camera_matrix)
    return calculate_symmetric_epipolar_distance(feature_a=feature_a, feature_b=feature_b, e=e)


def eight_point_model_fitter(matching_features: list, camera_matrix):
    """epipolar_ransac.py:28-42."""
    if 8 != len(matching_features):
        raise ValueError("Eight feature pairs are expected.")
    matches = [Match(a_index=i, b_index=i) for i in range(len(matching_features))]
    return estimate_essential_mat(
        camera_matrix=camera_matrix,
        features_a=[p[0] for p in matching_features],
        features_b=[p[1] for p in matching_features],
        matches=matches,
    )


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
