"""Shared plumbing of the patch scores: one GPU call for an [na, nb] block of feature pairs."""
from __future__ import annotations

import numpy as np

from .. import _native

MAX_WINDOW = 15  # csrc/sfm_match.cuh kMaxWindow


def check_images(image_a, image_b):
    image_a, image_b = np.asarray(image_a), np.asarray(image_b)
    if image_a.shape != image_b.shape:
        raise ValueError("the images must have the same shape")  # ncc.py:22-23, ssd.py:24-25
    if image_a.ndim != 2:
        raise ValueError("grayscale (2-D) images are expected")
    return image_a, image_b


def feature_array(features) -> np.ndarray:
    out = np.empty((len(features), 2), dtype=np.float64)
    for i, f in enumerate(features):
        out[i, 0] = f.x
        out[i, 1] = f.y
    return out


def single_score(kind, image_a, image_b, feature_a, feature_b, window_size):
    """One score through the same kernels as the matrix (na = nb = 1)."""
    image_a, image_b = check_images(image_a, image_b)
    eng = _native.get_engine()
    _, _, _, S = eng.match_brute_force(image_a, image_b, [[feature_a.x, feature_a.y]], [[feature_b.x, feature_b.y]],
                                       kind=kind, window=int(window_size), want_scores=True)
    return float(S[0, 0])
