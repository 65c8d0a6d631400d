"""Drop-in mirror of lib/epipolar/eight_point.py: eight-point estimation of the fundamental /
essential matrix, decomposition into rotation and translation, cheirality vote.

Public names, argument meaning, return types and exceptions follow the reference
(file:line given per function); the numerics run in libsfm_b200.so on the GPU.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .. import two_view
from ..common.feature import Feature
from ..errors import EightPointCalculationError  # noqa: F401  (re-exported, eight_point.py:20-23)
from ..feature_matching.matching import Match
from ..transforms.transforms import Transform3D
from .triangulation import triangulate_point_correspondence


def _coords(features) -> np.ndarray:
    return np.array([[f.x, f.y] for f in features], dtype=np.float64).reshape(-1, 2)


def estimate_r_t(camera_matrix, features_a: List[Feature], features_b: List[Feature], matches: List[Match]):
    """eight_point.py:26-62 — E from exactly eight matches, then (R, t, mask)."""
    if not features_a or not features_b:
        raise ValueError("Need some matching features")
    e = estimate_essential_mat(camera_matrix=camera_matrix, features_a=features_a, features_b=features_b,
                               matches=matches)
    a = [features_a[m.a_index] for m in matches]
    b = [features_b[m.b_index] for m in matches]
    return recover_r_t_from_e(e=e, camera_matrix=camera_matrix, features_a=a, features_b=b)


def recover_r_t_from_e(e, camera_matrix, features_a: list, features_b: list, distance_threshold=None):
    """eight_point.py:65-96 — (cam2_R_cam1, cam2_t_cam2_cam1, mask); mask is an int64 index array."""
    pa, pb = _coords(features_a), _coords(features_b)
    n = min(len(pa), len(pb))
    res = two_view.recover_pose_arrays(e, pa[:n], pb[:n], distance_threshold, camera_matrix=camera_matrix)
    return res.R, res.t, res.passing_indices


def estimate_essential_mat(*, camera_matrix, features_a: List[Feature], features_b: List[Feature],
                           matches: List[Match]):
    """eight_point.py:99-124."""
    if 8 != len(matches):
        raise ValueError("Exactly eight matches are needed")
    ca, cb = _get_matching_coordinates(features_a, features_b, matches)
    return two_view.eight_point_arrays(ca, cb, camera_matrix)


def to_normalized_image_coords(feature: Feature, camera_matrix) -> Feature:
    """eight_point.py:127-133 — skew and K[2,2] are ignored, as in the reference."""
    f_x = camera_matrix[0][0]
    f_y = camera_matrix[1][1]
    c_x = camera_matrix[0][2]
    c_y = camera_matrix[1][2]
    return Feature(x=(feature.x - c_x) / f_x, y=(feature.y - c_y) / f_y)


def estimate_fundamental_mat(features_a: List[Feature], features_b: List[Feature], matches: List[Match]) -> np.ndarray:
    """eight_point.py:136-170 — raises EightPointCalculationError on a degenerate sample."""
    if 8 != len(matches):
        raise ValueError("Exactly eight matches are needed")
    ca, cb = _get_matching_coordinates(features_a, features_b, matches)
    return two_view.eight_point_arrays(ca, cb, None)


def create_trivial_matches(num_features: int) -> list:
    """eight_point.py:173-178."""
    return [Match(a_index=i, b_index=i, match_score=0.0) for i in range(num_features)]


def _recover_r_t(features_a: list, features_b: list, e, distance_threshold=None):
    """eight_point.py:181-242 — inputs in normalised image coordinates."""
    n = min(len(features_a), len(features_b))
    res = two_view.recover_pose_arrays(e, _coords(features_a[:n]), _coords(features_b[:n]), distance_threshold)
    return res.R, res.t, res.passing_indices


def _recover_all_r_t(e) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """eight_point.py:245-280 — (R_1, R_2, t_1).  The (R_1, R_2) order and the sign of t_1
    depend on the SVD's sign convention (LAPACK there, Jacobi here); the candidate set is
    the same and the reference's test accepts either (tests/test_epipolar.py:205-229)."""
    return two_view.recover_all_r_t_arrays(e)


def _get_matching_coordinates(features_a, features_b, matches) -> Tuple[np.ndarray, np.ndarray]:
    """eight_point.py:283-305 — two Nx2 arrays of matched coordinates (a by a_index, b by b_index)."""
    ca = np.empty((len(matches), 2), dtype=np.float64)
    cb = np.empty((len(matches), 2), dtype=np.float64)
    for i, m in enumerate(matches):
        fa, fb = features_a[m.a_index], features_b[m.b_index]
        ca[i, 0], ca[i, 1] = fa.x, fa.y
        cb[i, 0], cb[i, 1] = fb.x, fb.y
    return ca, cb


def _normalize_coords(coords: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """eight_point.py:308-338 — Hartley normalisation (host helper kept for API parity; the
    fitter kernel does the same arithmetic per hypothesis on the device)."""
    centroid = np.mean(coords, axis=0)
    centered = coords - centroid
    scale = np.sqrt(2.0) / np.mean(np.linalg.norm(centered, axis=1))
    t = np.array([[scale, 0.0, -scale * centroid[0]], [0.0, scale, -scale * centroid[1]], [0.0, 0.0, 1.0]],
                 dtype=np.float64)
    return centered * scale, t


def _get_normalized_match_coordinates(features_a, features_b, matches):
    """eight_point.py:341-360."""
    ca, cb = _get_matching_coordinates(features_a, features_b, matches)
    return (*_normalize_coords(ca), *_normalize_coords(cb))


def _get_y_col(coord_a: np.ndarray, coord_b: np.ndarray):
    """eight_point.py:376-393 — one row of the design matrix (b-major: x_b^T F x_a = 0)."""
    assert 2 == len(coord_a)
    assert 2 == len(coord_b)
    return np.array([coord_b[0] * coord_a[0], coord_b[0] * coord_a[1], coord_b[0],
                     coord_b[1] * coord_a[0], coord_b[1] * coord_a[1], coord_b[1],
                     coord_a[0], coord_a[1], 1.0], dtype=np.float64)


def _cheirality_check(feature_a: Feature, feature_b: Feature, cam2_R_cam1, cam2_t_cam2_cam1,
                      distance_threshold=None, z_axis_index: int = 2) -> bool:
    """eight_point.py:449-488 — is the triangulated point in front of both cameras and near?"""
    if distance_threshold is None:
        distance_threshold = 50.0
    P1 = Transform3D.identity().Tmat
    P2 = Transform3D.from_rmat_t(np.asarray(cam2_R_cam1), np.asarray(cam2_t_cam2_cam1)).Tmat
    x1 = triangulate_point_correspondence(feature_a, feature_b, P1, P2)
    x2 = (P2 @ [*x1, 1])[:-1]
    TOLERANCE = 1e-8
    return bool(x1[z_axis_index] >= -TOLERANCE and x2[z_axis_index] >= -TOLERANCE
                and np.linalg.norm(x1) <= distance_threshold)
