"""Gaussian smoothing kernel with the API of lib/blur/gaussian.py:4-26.  A kernel_size x kernel_size table of
parameters built once on the host (a few dozen numbers); the filtering itself is ``cross_correlate`` on the GPU."""
import numpy as np


def create_gaussian_kernel(kernel_size: int, sigma: float) -> np.ndarray:
    """Normalised (sum 1) isotropic Gaussian sampled on the integer grid [-k, k]^2, k = kernel_size // 2.
    ValueError for sizes below 3 and for even sizes, as the reference (gaussian.py:13-16)."""
    if kernel_size <= 2:
        raise ValueError("kernel_size must be at least 3")
    if kernel_size % 2 == 0:
        raise ValueError("Only odd-sized kernels are accepted")
    offsets = np.arange(kernel_size) - kernel_size // 2
    squared_radius = np.add.outer(offsets ** 2, offsets ** 2)          # y^2 + x^2 on the grid (integers)
    weights = np.exp(-squared_radius / (2 * sigma ** 2))
    weights /= 2 * np.pi * sigma ** 2                                   # the reference applies the density constant
    weights /= np.sum(weights)                                          # before normalising: kept for bit parity
    return weights
