"""Mirror of lib/feature_matching/ncc.py:7-54 — the patch score apps/sfm.py:74 hands to the matcher."""
import numpy as np

from ..common import feature as feat
from . import _patch


def calculate_ncc(
    image_a: np.ndarray,
    image_b: np.ndarray,
    feature_a: feat.Feature,
    feature_b: feat.Feature,
    window_size: int = 3,
) -> float:
    """1 - NCC of the two windows, in [0, 2]; 2.0 when a window leaves the image (ncc.py:25-31) or has
    no texture (ncc.py:47-48).  Raises ValueError when the image shapes differ (ncc.py:22-23).

    One pair per call costs a kernel launch; ``match_brute_force`` recognises this function (bound with
    ``functools.partial`` or wrapped like apps/sfm.py:266-277) and scores all pairs in one launch."""
    return _patch.single_score("ncc", image_a, image_b, feature_a, feature_b, window_size)
