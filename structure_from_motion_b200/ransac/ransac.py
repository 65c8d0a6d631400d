"""Drop-in mirror of lib/ransac/ransac.py (ErrorAggregationMethod :12-16, fit_with_ransac :19-93).

``fit_with_ransac`` keeps the reference signature.  When the model is the eight-point
essential matrix scored by symmetric epipolar distance — i.e. the fitter/scorer are the
``functools.partial`` objects that lib/epipolar/epipolar_ransac.py:58-67 builds — the whole
loop runs on the GPU (sm_100a kernels behind the C ABI; no CPU fallback).  For arbitrary
Python callables (the reference's 2-point line model, lib/ransac/tests/test_ransac.py:103-111)
the callables themselves are host Python, so the driver loop necessarily runs on the host;
it follows ransac.py:61-86 step for step.
"""
from __future__ import annotations

import copy
import functools
import random
from enum import Enum
from math import inf
from typing import Any, Callable, Optional, Sequence, Tuple

import numpy as np


class ErrorAggregationMethod(Enum):
    SUM = "sum"
    SQUARE = "square"
    MEAN = "mean"
    RMS = "rms"


def _epipolar_camera_matrix(model_fit_data_count, model_fitter, inlier_scorer):
    """Return K if (fitter, scorer) are the epipolar pair of epipolar_ransac.py:58-67, else None."""
    from ..epipolar import epipolar_ransac as er

    if model_fit_data_count != 8:
        return None
    if not (isinstance(model_fitter, functools.partial) and isinstance(inlier_scorer, functools.partial)):
        return None
    if model_fitter.func is not er.eight_point_model_fitter or inlier_scorer.func is not er.calculate_sed_inlier_score:
        return None
    if model_fitter.args or inlier_scorer.args:
        return None
    ka = model_fitter.keywords.get("camera_matrix")
    kb = inlier_scorer.keywords.get("camera_matrix")
    if ka is None or kb is None or set(model_fitter.keywords) != {"camera_matrix"} \
            or set(inlier_scorer.keywords) != {"camera_matrix"}:
        return None
    if not np.array_equal(np.asarray(ka), np.asarray(kb)):
        return None
    return np.asarray(ka, dtype=np.float64)


def fit_with_ransac(
    data: Sequence,
    model_fit_data_count: int,
    model_fitter: Callable[[Sequence], Any],
    inlier_scorer: Callable[[Any, Any], float],
    inlier_threshold: float,
    min_num_extra_inliers: int | None = None,
    error_aggregation_method: ErrorAggregationMethod | None = None,
    max_iterations: int | None = None,
) -> Tuple[Optional[Any], Sequence]:
    """Fit a model using RANSAC; uses (and advances) the built-in ``random`` module's global
    state exactly as the reference does.  Returns (best model, its inliers)."""
    K = _epipolar_camera_matrix(model_fit_data_count, model_fitter, inlier_scorer)
    if K is not None:
        from .. import two_view

        pts_a = np.array([[p[0].x, p[0].y] for p in data], dtype=np.float64).reshape(-1, 2)
        pts_b = np.array([[p[1].x, p[1].y] for p in data], dtype=np.float64).reshape(-1, 2)
        res = two_view.ransac_essential_arrays(
            K, pts_a, pts_b, inlier_threshold, min_num_extra_inliers, error_aggregation_method,
            max_iterations, sampler="reference")
        # ransac.py:59 works on a deep copy; the returned inliers are copies, never the caller's objects
        return res.E, [copy.deepcopy(data[int(i)]) for i in res.inlier_indices]
    return _fit_with_ransac_callables(data, model_fit_data_count, model_fitter, inlier_scorer,
                                      inlier_threshold, min_num_extra_inliers, error_aggregation_method,
                                      max_iterations)


def _fit_with_ransac_callables(data, k, model_fitter, inlier_scorer, inlier_threshold,
                               min_num_extra_inliers, error_aggregation_method, max_iterations):
    """Generic driver for user-supplied Python callables (ransac.py:48-93)."""
    if max_iterations is None:
        max_iterations = 100
    if error_aggregation_method is None:
        error_aggregation_method = ErrorAggregationMethod.RMS
    if min_num_extra_inliers is None:
        min_num_extra_inliers = 0
    best_model, best_inliers, best_error = None, [], inf
    pool = copy.deepcopy(data)
    for _ in range(max_iterations):
        random.shuffle(pool)
        head, tail = pool[:k], pool[k:]
        model = model_fitter(head)
        extra = [d for d in tail if inlier_scorer(model, d) <= inlier_threshold]
        if min_num_extra_inliers <= len(extra):
            chosen = head + extra
            error = _aggregate_error([inlier_scorer(model, d) for d in chosen], error_aggregation_method)
            if error < best_error:
                best_model, best_inliers, best_error = model, chosen, error
    if best_model is None:
        raise ValueError(f"No model could be found with at least {min_num_extra_inliers + k} inliers.")
    return best_model, best_inliers


def _aggregate_error(errors: list, aggregation_method: ErrorAggregationMethod) -> float:
    """ransac.py:96-108."""
    v = aggregation_method.value
    if v == ErrorAggregationMethod.SUM.value:
        return sum(errors)
    if v == ErrorAggregationMethod.SQUARE.value:
        return np.sum(np.square(errors)).item()
    if v == ErrorAggregationMethod.MEAN.value:
        return np.mean(errors).item()
    if v == ErrorAggregationMethod.RMS.value:
        return np.sqrt(np.mean(np.square(errors))).item()
    raise NotImplementedError(aggregation_method)
