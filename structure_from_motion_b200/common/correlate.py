"""Drop-in mirror of lib/common/correlate.py:4-39 (``sfm_cross_correlate``; no CPU fallback)."""
import numpy as np


def cross_correlate(image: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """Slide the kernel over the image: per-pixel cross correlation with zero "same" padding.
    Only odd-sized square kernels; the same ValueErrors as the reference (correlate.py:13-22)."""
    from .. import _native
    from ..harris.harris_detector import _image_for_device

    image, kernel = np.asarray(image), np.asarray(kernel)
    if len(image.shape) != 2 or len(kernel.shape) != 2:
        raise ValueError("Only 2D single channel images are supported")
    if kernel.shape[0] != kernel.shape[1] or (kernel.shape[0] % 2) == 0:
        raise ValueError("Only odd-sized square kernels are supported")
    if image.shape[0] < kernel.shape[0] or image.shape[1] < kernel.shape[0]:
        raise ValueError("Kernel cannot be larger than image")
    return _native.get_engine().cross_correlate(_image_for_device(image), kernel)
