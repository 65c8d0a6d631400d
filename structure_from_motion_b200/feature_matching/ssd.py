"""Mirror of lib/feature_matching/ssd.py:7-36."""
import numpy as np

from ..common import feature as feat
from . import _patch


def calculate_ssd(
    image_a: np.ndarray,
    image_b: np.ndarray,
    feature_a: feat.Feature,
    feature_b: feat.Feature,
    window_size: int = 5,
) -> float:
    """Mean squared difference of the two windows; inf when a window leaves the image (ssd.py:27-30).
    uint8 images follow numpy's uint8 arithmetic (difference and square wrap modulo 256) exactly as
    the reference does when handed cv.cvtColor output.  ValueError when the shapes differ (ssd.py:24-25)."""
    return _patch.single_score("ssd", image_a, image_b, feature_a, feature_b, window_size)
