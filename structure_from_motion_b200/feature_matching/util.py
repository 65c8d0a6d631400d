"""Window bookkeeping of the patch scores — the two helpers of lib/feature_matching/util.py:8-27, re-implemented.
No arithmetic happens here; the CUDA kernel ``k_patch_prepare`` applies the same two rules on the device."""
from typing import Tuple

import numpy as np

from ..common import feature as feat


def _reach(window_size: int) -> int:
    """Pixels covered on each side of the centre: int(window_size / 2), so an even size w spans w + 1 pixels."""
    return int(window_size / 2)


def is_within_bounds(feature: feat.Feature, image_shape: Tuple[int, int], window_size: int) -> bool:
    """True iff the window around the feature stays inside the image.  The FLOAT coordinates are compared
    (util.py:12-16), while the window itself is cut at the truncated coordinates."""
    reach = _reach(window_size)
    rows, cols = image_shape[0], image_shape[1]
    return bool(reach <= feature.y < rows - reach and reach <= feature.x < cols - reach)


def select_window(image: np.ndarray, feature: feat.Feature, window_size: int) -> np.ndarray:
    """The (2 reach + 1)^2 view centred on (int(y), int(x)) (util.py:21-27)."""
    reach = _reach(window_size)
    row, col = int(feature.y), int(feature.x)
    return image[row - reach: row + reach + 1, col - reach: col + reach + 1]
