"""Drop-in mirror of lib/harris/harris_detector.py:11-55 — the first stage of apps/sfm.py:64-71
(SURVEY.md §8(f) N2).  The reference loops over the pixels in Python; here the Sobel responses, the
cornerness, the (scan-order dependent) non-maximum suppression and the top-``num_corners`` selection are
CUDA kernels (csrc/sfm_harris.cuh behind ``sfm_harris_corners``).  No CPU fallback."""
from __future__ import annotations

from typing import List

import numpy as np

from ..common import feature


def _image_for_device(image: np.ndarray):
    image = np.asarray(image)
    if image.ndim != 2:
        raise ValueError("Only 2D single channel images are supported")  # correlate.py:13-14
    if image.dtype == np.uint8:
        return np.ascontiguousarray(image)
    return np.ascontiguousarray(image, dtype=np.float64)


def detect_harris_corners(
    image: np.ndarray, num_corners: int = 50, block_size: int = 2, k: float = 0.04
) -> List[feature.Feature]:
    """Corners with the highest cornerness ``det(M) - k trace(M)^2`` (M = block_size x block_size sums of the
    Sobel products), at most ``num_corners``, in descending order of cornerness, as ``Feature(x, y)`` with
    ``x = column + block_size / 2``.  Pixels with zero cornerness are never returned.  Among corners of exactly
    equal cornerness the order is descending flat index (numpy leaves it unspecified)."""
    from .. import _native

    if num_corners <= 0:
        raise ValueError("num_corners needs to be at least 1")
    xy, _, _ = _native.get_engine().harris_corners(_image_for_device(image), int(num_corners), int(block_size), float(k))
    return [feature.Feature(x=np.float64(x), y=np.float64(y)) for x, y in xy]
