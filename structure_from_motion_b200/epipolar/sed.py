"""Drop-in mirror of lib/epipolar/sed.py:7-30."""
from __future__ import annotations

import numpy as np

from ..common.feature import Feature


def calculate_symmetric_epipolar_distance(feature_a: Feature, feature_b: Feature, e) -> float:
    """Symmetric Epipolar Distance (Hartley & Zisserman 11.10) of one correspondence in
    normalised image coordinates under the essential matrix ``e``; evaluated by the device
    scorer (the same routine the RANSAC kernels use), bit-compatible with the reference's
    numpy evaluation order."""
    from .. import two_view

    sed = two_view.sed_arrays(e, np.array([[feature_a.x, feature_a.y]], dtype=np.float64),
                              np.array([[feature_b.x, feature_b.y]], dtype=np.float64))
    return float(sed[0])
