"""Drop-in mirror of lib/epipolar/triangulation.py (triangulate_point_correspondence :9-39,
triangulate_points :42-62)."""
from __future__ import annotations

import numpy as np

from ..common.feature import Feature
from ..transforms.transforms import Transform3D


def _coords(features) -> np.ndarray:
    return np.array([[f.x, f.y] for f in features], dtype=np.float64).reshape(-1, 2)


def triangulate_point_correspondence(feature_a: Feature, feature_b: Feature, P1, P2) -> np.ndarray:
    """3D position of the point seen as feature_a by camera P1 and feature_b by camera P2
    (3x4 or 4x4 camera matrices; rows 0-2 are used, as in the reference)."""
    from .. import two_view

    X = two_view.triangulate_arrays(_coords([feature_a]), _coords([feature_b]), np.asarray(P1), np.asarray(P2))
    return X[0]


def triangulate_points(features_a, features_b, intrinsic_camera_matrix, cam2_T_cam1: Transform3D) -> np.ndarray:
    from .. import two_view

    P1, P2 = two_view.camera_matrices(intrinsic_camera_matrix, cam2_T_cam1.Tmat)
    if len(features_a) == 0:
        return np.array([])
    n = min(len(features_a), len(features_b))  # zip semantics of triangulation.py:57-62
    return two_view.triangulate_arrays(_coords(features_a[:n]), _coords(features_b[:n]), P1, P2)
