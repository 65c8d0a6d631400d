"""Mirror of lib/blur/gaussian.py:4-26: a kernel_size x kernel_size table of parameters (a few dozen
numbers built once on the host; the filtering itself is ``cross_correlate`` on the GPU)."""
import numpy as np


def create_gaussian_kernel(kernel_size: int, sigma: float) -> np.ndarray:
    """Normalised Gaussian smoothing kernel; ValueError for sizes <= 2 or even sizes (gaussian.py:13-16)."""
    if kernel_size <= 2:
        raise ValueError("kernel_size must be at least 3")
    if kernel_size % 2 == 0:
        raise ValueError("Only odd-sized kernels are accepted")
    half = int(kernel_size / 2)
    coords = np.arange(-half, half + 1)
    x_grid, y_grid = np.meshgrid(coords, coords)
    kernel = np.exp(-(x_grid ** 2 + y_grid ** 2) / (2 * sigma ** 2))
    kernel /= 2 * np.pi * sigma ** 2
    kernel /= np.sum(kernel)
    return kernel
