"""Array-native entry points of the two-view hot path (numpy in, numpy out).

These are the twins of the reference's list-of-dataclass callables:
marshalling 10^5..10^6 ``Feature`` objects costs more than the GPU work, so the
list-based mirrors in ``ransac/`` and ``epipolar/`` convert once and call these.
Everything numeric happens in libsfm_b200.so (CUDA, sm_100a); there is no CPU fallback.
"""
from __future__ import annotations

import dataclasses
import random
from math import inf
from typing import Optional

import numpy as np

from . import _native
from .errors import EIGHT_POINT_ERROR_MESSAGE, EightPointCalculationError
from .ransac.ransac import ErrorAggregationMethod


@dataclasses.dataclass
class RansacResult:
    E: np.ndarray                 # (3,3), E[2,2] == 1
    best_index: int               # winning iteration (global hypothesis index)
    error: float                  # aggregated error of the winner
    count_extra: int              # inliers beyond the 8 samples
    inlier_indices: np.ndarray    # int64; reference order (samples, then permuted rest) with the
                                  # reference sampler, else samples then ascending
    mask: np.ndarray              # bool[n]: sed <= threshold under the winner
    sed: np.ndarray               # float64[n]
    sample: np.ndarray            # int32[8] the winner's minimal sample
    num_invalid: int
    first_invalid: int


def _agg_name(method) -> str:
    if method is None:
        return "rms"  # ransac.py:50-51
    # the reference compares ``.value`` (ransac.py:99-105), so any Enum with the same values is accepted
    return str(getattr(method, "value", method))


def _reference_error(errors, agg: str) -> float:
    """_aggregate_error (lib/ransac/ransac.py:96-108) on a python list in [samples..., extra inliers...] order:
    the reference's own summation order (python ``sum`` / numpy pairwise) — used only to settle near-ties."""
    if agg == "sum":
        return sum(errors)
    if agg == "square":
        return np.sum(np.square(errors)).item()
    if agg == "mean":
        return np.mean(errors).item()
    return np.sqrt(np.mean(np.square(errors))).item()


def _resolve_near_ties(eng, ties, threshold, agg, rows, order_after):
    """SURVEY.md H1.  The GPU selects on exactly rounded sums; the reference keeps ``model_error < best_model_error``
    (ransac.py:83) on errors summed in list order.  For the hypotheses whose errors agree to 1e-12 the list-order error
    is recomputed here from the device's exact per-correspondence scores and the strict ``<`` is replayed in iteration
    order.  rows(t) -> the 8 sample indices of hypothesis t; order_after(t) -> the correspondences that follow the
    samples in that iteration's list (the permutation tail for the reference sampler).  Returns the local index."""
    best_t, best_err = None, float("inf")
    for t in ties:  # ascending = iteration order
        eng.set_winner(int(t))
        _, sed = eng.inlier_mask(threshold)
        rest = order_after(int(t))
        errs = [float(v) for v in sed[rows(int(t))]] + [float(v) for v in sed[rest][sed[rest] <= threshold]]
        e = _reference_error(errs, agg)
        if e < best_err:
            best_t, best_err = int(t), e
    return best_t, best_err


def _get_state_words():
    version, words, gauss = random.getstate()
    return version, np.array(words, dtype=np.uint32), gauss


def _set_state_words(version, words, gauss):
    random.setstate((version, tuple(int(w) for w in words), gauss))


def ransac_essential_arrays(
    camera_matrix,
    pts_a,
    pts_b,
    threshold: float,
    min_num_extra_inliers=None,
    error_aggregation_method=None,
    max_iterations: Optional[int] = None,
    *,
    sampler: str = "reference",
    seed: int = 0,
    hyp_offset: int = 0,
    on_degenerate: str = "raise",
    selection: str = "min_error",
    engine: Optional[_native.Engine] = None,
    table: Optional[np.ndarray] = None,
    _tail: Optional[dict] = None,
) -> RansacResult:
    """estimate_essential_mat_with_ransac (lib/epipolar/epipolar_ransac.py:45-70) on arrays.

    pts_a, pts_b: [n,2] matched pixel coordinates (row i of each is one correspondence).
    sampler: "reference" — CPython random.shuffle semantics on the process-global state
             (lib/ransac/ransac.py:62), the state advances exactly as in the reference;
             "device" — Philox sampler on the GPU (seed), for sizes where the reference's
             O(N) shuffle per iteration is infeasible; "table" — use the given table.
    """
    if max_iterations is None:
        max_iterations = 100  # ransac.py:48-49
    if min_num_extra_inliers is None:
        min_num_extra_inliers = 0  # ransac.py:52-53
    agg = _agg_name(error_aggregation_method)
    pts_a = np.ascontiguousarray(pts_a, dtype=np.float64).reshape(-1, 2)
    pts_b = np.ascontiguousarray(pts_b, dtype=np.float64).reshape(-1, 2)
    n = pts_a.shape[0]
    H = int(max_iterations)

    def no_model():
        return ValueError(f"No model could be found with at least {min_num_extra_inliers + 8} inliers.")

    if H <= 0:
        raise no_model()  # ransac.py:88-91 (loop never ran)
    if n < 8:
        if sampler == "reference":
            random.shuffle(list(range(n)))  # the reference shuffles once before the fitter raises
        raise ValueError("Eight feature pairs are expected.")  # epipolar_ransac.py:31-32
    eng = engine or _native.get_engine()

    if sampler == "reference":
        version, words, gauss = _get_state_words()
        ref_sampler = _native.ReferenceSampler(words, n, H)
        table, words = ref_sampler.table, ref_sampler.final_state
    elif sampler == "table":
        table = np.ascontiguousarray(table, dtype=np.int32).reshape(-1, 8)
        H = table.shape[0]

    eng.upload_pairs(pts_a, pts_b, camera_matrix)
    if sampler == "device":
        eng.sample_device(seed, H, hyp_offset=hyp_offset)
    else:
        eng.set_table(table)
    if _tail is None:
        best, mask, sed = eng.ransac_essential(threshold, float(min_num_extra_inliers), agg, selection)
    else:  # two_view_arrays: selection, inlier mask, pose vote and triangulation in one C call
        best, mask, sed, *rest = eng.two_view(threshold, float(min_num_extra_inliers), agg, selection,
                                              _tail["distance_threshold"])
        _tail["out"] = rest

    if sampler == "reference":
        if best.num_invalid > 0 and on_degenerate == "raise":
            # the reference aborts inside iteration first_invalid: only that many shuffles happened
            w, _ = ref_sampler.after(int(best.first_invalid))
            _set_state_words(version, w, gauss)
        else:
            _set_state_words(version, words, gauss)
    if best.num_invalid > 0 and on_degenerate == "raise":
        raise EightPointCalculationError(EIGHT_POINT_ERROR_MESSAGE)  # eight_point.py:417-421
    if best.index < 0:
        raise no_model()

    if sampler != "device" and selection == "min_error":
        # H1: a definite list order exists (reference sampler: the permutation; table: samples then ascending), so
        # near-ties are settled in the reference's own summation order
        ties, total = eng.near_ties(1e-12, 64)
        if total > 1 and total <= 64:
            def rows(t):
                return np.asarray(table[t], dtype=np.int64)

            def order_after(t):
                if sampler == "reference":
                    return ref_sampler.after(t)[1][8:].astype(np.int64)
                keep = np.ones(n, dtype=bool)
                keep[table[t]] = False
                return np.nonzero(keep)[0]

            t, _ = _resolve_near_ties(eng, ties, threshold, agg, rows, order_after)
            if t is not None and t != int(best.index):
                E_t, _ = eng.get_models(t, 1)
                best.index = t
                best.E[:] = tuple(E_t.reshape(9))
                eng.set_winner(t)
                m2, sed = eng.inlier_mask(threshold)
                mask = m2.astype(np.uint8)
                best.count_extra = int(m2.sum() - m2[table[t]].sum())
                errs = [float(v) for v in sed[rows(t)]] + [float(v) for v in sed[order_after(t)][sed[order_after(t)] <= threshold]]
                best.err = _reference_error(errs, agg)
                if _tail is not None:
                    _tail["out"] = list(eng.pose_and_triangulate(threshold, _tail["distance_threshold"]))
            else:
                eng.set_winner(int(best.index))

    mask = mask.astype(bool)
    local = int(best.index)
    sample = np.array(best.sample, dtype=np.int32) if sampler == "device" else table[local]
    is_sample = np.zeros(n, dtype=bool)
    is_sample[sample] = True
    if sampler == "reference":
        # inliers in the reference's order: samples, then the rest of that iteration's permutation
        _, perm = ref_sampler.after(int(best.index))  # replayed from the nearest snapshot
        rest = perm[8:].astype(np.int64)
        extra = rest[mask[rest]]
    else:
        extra = np.nonzero(mask & ~is_sample)[0]
    inliers = np.concatenate([np.asarray(sample, dtype=np.int64), extra])
    return RansacResult(
        E=np.array(best.E, dtype=np.float64).reshape(3, 3),
        best_index=int(best.index) + (hyp_offset if sampler == "device" else 0),
        error=float(best.err),
        count_extra=int(best.count_extra),
        inlier_indices=inliers,
        mask=mask,
        sed=sed,
        sample=np.asarray(sample, dtype=np.int32),
        num_invalid=int(best.num_invalid),
        first_invalid=int(best.first_invalid),
    )


def ransac_essential_adaptive(camera_matrix, pts_a, pts_b, threshold, min_num_extra_inliers=None,
                              error_aggregation_method=None, max_iterations: int = 65536, *, confidence: float = 0.99,
                              chunk: int = 4096, seed: int = 0, selection: str = "max_inliers",
                              engine: Optional[_native.Engine] = None):
    """Early termination (the reference always runs ``max_iterations``): hypotheses are drawn
    by the device sampler and evaluated ``chunk`` at a time; after each chunk the standard RANSAC bound
    ``log(1 - confidence) / log(1 - w^8)`` is recomputed from the inlier ratio ``w`` of the best model so far and the
    run stops once that many hypotheses have been evaluated.  Because the sampler is keyed by the global hypothesis
    index, the result equals ``ransac_essential_arrays(..., sampler="device")`` over the same number of hypotheses.
    Returns (RansacResult, hypotheses evaluated)."""
    from .distributed import merge_best, pack_local_best

    eng = engine or _native.get_engine()
    n = np.asarray(pts_a).reshape(-1, 2).shape[0]
    best_row, best_res, done = None, None, 0
    needed = float(max_iterations)
    while done < max_iterations and done < needed:
        h = int(min(chunk, max_iterations - done))
        try:
            res = ransac_essential_arrays(camera_matrix, pts_a, pts_b, threshold, min_num_extra_inliers,
                                          error_aggregation_method, h, sampler="device", seed=seed, hyp_offset=done,
                                          on_degenerate="skip", selection=selection, engine=eng)
            row = pack_local_best(res.error, res.best_index, res.count_extra, res.E)
            rows = row[None] if best_row is None else np.stack([best_row, row])
            owner, *_ = merge_best(rows, selection)
            if best_row is None or owner == 1:
                best_row, best_res = row, res
        except ValueError:
            pass  # no candidate in this chunk
        done += h
        if best_res is not None:
            w = min(1.0, (8 + best_res.count_extra) / max(n, 1))
            miss = 1.0 - w ** 8
            needed = 0.0 if miss <= 0.0 else (np.log(1.0 - confidence) / np.log(miss) if miss < 1.0 else float(max_iterations))
    if best_res is None:
        raise ValueError(f"No model could be found with at least {(min_num_extra_inliers or 0) + 8} inliers.")
    return best_res, done


def eight_point_arrays(pts_a, pts_b, camera_matrix=None, engine=None) -> np.ndarray:
    """estimate_essential_mat / estimate_fundamental_mat on 8 gathered pairs ([8,2] arrays).

    lib/epipolar/eight_point.py:99-170; camera_matrix None = fundamental matrix (pixel coords).
    """
    pts_a = np.ascontiguousarray(pts_a, dtype=np.float64).reshape(-1, 2)
    pts_b = np.ascontiguousarray(pts_b, dtype=np.float64).reshape(-1, 2)
    if pts_a.shape[0] != 8 or pts_b.shape[0] != 8:
        raise ValueError("Exactly eight matches are needed")  # eight_point.py:151-152
    eng = engine or _native.get_engine()
    K = np.eye(3) if camera_matrix is None else camera_matrix
    eng.upload_pairs(pts_a, pts_b, K)
    eng.set_table(np.arange(8, dtype=np.int32).reshape(1, 8))
    E, valid, _ = eng.fit()
    if not valid[0]:
        raise EightPointCalculationError(EIGHT_POINT_ERROR_MESSAGE)
    return E[0]


def sed_arrays(e, norm_a, norm_b, engine=None) -> np.ndarray:
    """calculate_symmetric_epipolar_distance (lib/epipolar/sed.py:7-30) for [n,2] arrays of
    K-normalised coordinates and one E."""
    eng = engine or _native.get_engine()
    na = np.ascontiguousarray(norm_a, dtype=np.float64).reshape(-1, 2)
    nb = np.ascontiguousarray(norm_b, dtype=np.float64).reshape(-1, 2)
    eng.upload_pairs(na, nb, np.eye(3))
    eng.set_models(np.asarray(e, dtype=np.float64).reshape(1, 9))
    eng.set_winner(0)
    _, sed = eng.inlier_mask(inf)
    return sed


@dataclasses.dataclass
class PoseResult:
    R: np.ndarray
    t: np.ndarray
    passing_indices: np.ndarray  # int64 indices of correspondences passing the chosen pose
    counts: np.ndarray           # votes per candidate (with the reference's index-0 quirk)
    candidates: list             # [(R, t)] * 4 in the reference's enumeration order
    singular_values: np.ndarray


def _check_decomposition(p):
    # lib/epipolar/eight_point.py:268-271 — np.isclose(0.0, s[-1]) with default tolerances
    if not np.isclose(0.0, p.sv[2]):
        raise EightPointCalculationError(
            "The smallest singular value of the Essential matrix is expected to be ~0")


def recover_all_r_t_arrays(e, engine=None):
    """_recover_all_r_t (lib/epipolar/eight_point.py:245-280) -> (R1, R2, t1)."""
    eng = engine or _native.get_engine()
    p = eng.decompose_essential(e)
    _check_decomposition(p)
    R = np.array(p.R, dtype=np.float64).reshape(4, 3, 3)
    t = np.array(p.t, dtype=np.float64).reshape(4, 3)
    return R[0], R[2], t[0]


def recover_pose_arrays(e, pts_a, pts_b, distance_threshold=None, engine=None, camera_matrix=None) -> PoseResult:
    """_recover_r_t (lib/epipolar/eight_point.py:181-242) on [m,2] arrays of K-normalised coordinates; with
    ``camera_matrix`` the arrays hold pixel coordinates and to_normalized_image_coords (eight_point.py:127-133) runs
    on the device first — recover_r_t_from_e (eight_point.py:65-96)."""
    if distance_threshold is None:
        distance_threshold = 50.0  # eight_point.py:469-470
    eng = engine or _native.get_engine()
    na = np.ascontiguousarray(pts_a, dtype=np.float64).reshape(-1, 2)
    nb = np.ascontiguousarray(pts_b, dtype=np.float64).reshape(-1, 2)
    p, pass4 = eng.recover_pose(e, na, nb, distance_threshold, camera_matrix=camera_matrix)
    _check_decomposition(p)
    counts = np.array(p.counts, dtype=np.int64)
    if 0 == np.count_nonzero(counts):
        raise EightPointCalculationError("None of the transformations pass the cheirality check.")
    R = np.array(p.R, dtype=np.float64).reshape(4, 3, 3)
    t = np.array(p.t, dtype=np.float64).reshape(4, 3)
    b = int(p.best)
    idx = np.nonzero((pass4 >> b) & 1)[0].astype(np.int64)
    return PoseResult(R=R[b].copy(), t=t[b].copy(), passing_indices=idx, counts=counts,
                      candidates=[(R[i], t[i]) for i in range(4)], singular_values=np.array(p.sv))


def triangulate_arrays(pts_a, pts_b, P1, P2, engine=None) -> np.ndarray:
    """triangulate_point_correspondence over arrays (lib/epipolar/triangulation.py:9-39)."""
    eng = engine or _native.get_engine()
    return eng.triangulate(P1, P2, pts_a, pts_b)


def camera_matrices(intrinsic_camera_matrix, cam2_T_cam1_Tmat):
    """P1 = K[I|0], P2 = K[R|t] as lib/epipolar/triangulation.py:52-56 builds them."""
    K = np.asarray(intrinsic_camera_matrix, dtype=np.float64)
    if (3, 3) != K.shape:
        raise ValueError(f"Camera intrinsic matrix is not 3x3, actual shape: {K.shape}")
    K_ext = np.hstack((K, np.zeros((3, 1))))
    P1 = K_ext @ np.eye(4)
    P2 = K_ext @ (np.asarray(cam2_T_cam1_Tmat, dtype=np.float64) @ np.eye(4))
    return P1, P2


@dataclasses.dataclass
class TwoViewResult:
    ransac: RansacResult
    R: np.ndarray
    t: np.ndarray
    inlier_indices: np.ndarray   # ascending indices of the winner's inliers (samples included)
    passing: np.ndarray          # bool per inlier: passed the cheirality vote
    points: np.ndarray           # [m,3] triangulated points (NaN where not passing)
    counts: np.ndarray


def two_view_arrays(camera_matrix, pts_a, pts_b, threshold, min_num_extra_inliers=None,
                    error_aggregation_method=None, max_iterations=None, distance_threshold=None,
                    **kw) -> TwoViewResult:
    """The whole hot path of apps/sfm.py:110-186 in one call: RANSAC E -> cheirality vote ->
    triangulation of the passing inliers.  Correspondences are uploaded once; only results
    come back."""
    if distance_threshold is None:
        distance_threshold = 50.0
    eng = kw.get("engine") or _native.get_engine()
    kw["engine"] = eng
    if (kw.get("sampler") == "device" and kw.get("on_degenerate", "raise") == "skip" and kw.get("hyp_offset", 0) == 0
            and np.asarray(pts_a).reshape(-1, 2).shape[0] >= 8 and (max_iterations or 100) > 0):
        # device sampler, degenerate samples skipped: nothing on the host depends on intermediate results, so the
        # whole estimate is enqueued in one go (un-synchronised upload included) and fetched once
        return _two_view_device(eng, camera_matrix, pts_a, pts_b, threshold, min_num_extra_inliers or 0,
                                _agg_name(error_aggregation_method), int(max_iterations or 100), float(distance_threshold),
                                int(kw.get("seed", 0)), kw.get("selection", "min_error"))
    tail = {"distance_threshold": float(distance_threshold)}
    res = ransac_essential_arrays(camera_matrix, pts_a, pts_b, threshold, min_num_extra_inliers,
                                  error_aggregation_method, max_iterations, _tail=tail, **kw)
    p, num, idx, ok, X = tail["out"]
    _check_decomposition(p)
    counts = np.array(p.counts, dtype=np.int64)
    if 0 == np.count_nonzero(counts):
        raise EightPointCalculationError("None of the transformations pass the cheirality check.")
    b = int(p.best)
    R = np.array(p.R, dtype=np.float64).reshape(4, 3, 3)[b].copy()
    t = np.array(p.t, dtype=np.float64).reshape(4, 3)[b].copy()
    return TwoViewResult(ransac=res, R=R, t=t, inlier_indices=idx, passing=((ok >> b) & 1).astype(bool),
                         points=X, counts=counts)


@dataclasses.dataclass
class ImagePairResult:
    corners_a: np.ndarray        # [na,2] (x, y), descending cornerness
    corners_b: np.ndarray
    match_a: np.ndarray          # int64 indices into corners_a of the matches handed to RANSAC
    match_b: np.ndarray
    match_score: np.ndarray
    two_view: TwoViewResult      # indices inside it refer to the match arrays


def image_pair_arrays(image_a, image_b, camera_matrix, *, num_harris_corners: int = 600, ncc_window_size: int = 9,
                      ratio_test_threshold: float = 0.7, match_score_threshold: float = 0.3,
                      sed_inlier_threshold: float = 1.5e-6, min_num_extra_inliers=10, max_iterations: int = 2000,
                      error_aggregation_method="rms", sampler: str = "reference", seed: int = 0,
                      engine: Optional[_native.Engine] = None) -> ImagePairResult:
    """The whole of apps/sfm.py:64-186 on arrays, with the defaults of apps/config/config.yaml: Harris corners on
    both images -> brute-force NCC matching with ratio test + cross-check -> score filter (apps/sfm.py:280-296) ->
    RANSAC essential matrix -> cheirality vote -> triangulation.  No ``Feature`` / ``Match`` objects are built; with
    ``sampler="reference"`` the results equal those of the list-based pipeline (same global ``random`` stream)."""
    eng = engine or _native.get_engine()
    ca, _, _ = eng.harris_corners(image_a, num_harris_corners)
    cb, _, _ = eng.harris_corners(image_b, num_harris_corners)
    if len(ca) == 0 or len(cb) == 0:
        raise ValueError("Eight feature pairs are expected.")
    best_b, best_s, keep, _ = eng.match_brute_force(image_a, image_b, ca, cb, kind="ncc", window=ncc_window_size,
                                                    ratio_test=True, crosscheck=True, ratio_threshold=ratio_test_threshold)
    keep &= ~(best_s > match_score_threshold)
    ia = np.flatnonzero(keep)
    ib = best_b[ia].astype(np.int64)
    tv = two_view_arrays(camera_matrix, ca[ia], cb[ib], sed_inlier_threshold, min_num_extra_inliers,
                         error_aggregation_method, max_iterations, sampler=sampler, seed=seed, engine=eng)
    return ImagePairResult(corners_a=ca, corners_b=cb, match_a=ia, match_b=ib, match_score=best_s[ia], two_view=tv)


def _device_submit(eng, out, camera_matrix, pts_a, pts_b, threshold, min_extra, agg, max_iterations, distance_threshold,
                   seed, selection):
    eng.upload_pairs(pts_a, pts_b, camera_matrix, sync=False)
    eng.sample_device(seed, max_iterations)
    return eng.two_view_async(threshold, float(min_extra), agg, selection, distance_threshold, out=out)


def _device_result(eng, mask, sed, min_extra, copy: bool) -> "TwoViewResult":
    best, p, num, idx, ok, X = eng.two_view_fetch()
    if best.index < 0:
        raise ValueError(f"No model could be found with at least {min_extra + 8} inliers.")
    _check_decomposition(p)
    counts = np.array(p.counts, dtype=np.int64)
    if 0 == np.count_nonzero(counts):
        raise EightPointCalculationError("None of the transformations pass the cheirality check.")
    sample = np.array(best.sample, dtype=np.int32)
    extra = idx[~np.isin(idx, sample)]  # the tail's compacted list = sed <= thr plus the samples, ascending
    b = int(p.best)
    res = RansacResult(E=np.array(best.E, dtype=np.float64).reshape(3, 3), best_index=int(best.index),
                       error=float(best.err), count_extra=int(best.count_extra),
                       inlier_indices=np.concatenate([sample.astype(np.int64), extra]), mask=mask.astype(bool),
                       sed=sed.copy() if copy else sed, sample=sample, num_invalid=int(best.num_invalid),
                       first_invalid=int(best.first_invalid))
    R = np.array(p.R, dtype=np.float64).reshape(4, 3, 3)[b].copy()
    t = np.array(p.t, dtype=np.float64).reshape(4, 3)[b].copy()
    return TwoViewResult(ransac=res, R=R, t=t, inlier_indices=idx, passing=((ok >> b) & 1).astype(bool), points=X,
                         counts=counts)


def _two_view_device(eng, camera_matrix, pts_a, pts_b, threshold, min_extra, agg, max_iterations, distance_threshold, seed,
                     selection):
    n = np.asarray(pts_a).reshape(-1, 2).shape[0]
    out = eng.pinned_out(n)  # engine-owned pinned landing buffers: page-locking per call would dominate the transfers
    mask, sed = _device_submit(eng, out, camera_matrix, pts_a, pts_b, threshold, min_extra, agg, max_iterations,
                               distance_threshold, seed, selection)
    return _device_result(eng, mask, sed, min_extra, copy=True)


class TwoViewStream:
    """Back-to-back estimates from HOST buffers with the transfers of one estimate hidden behind the kernels of
    another: ``depth`` contexts on the same GPU (each its own stream); ``submit`` only ENQUEUES the upload of the
    correspondences, the device sampler, fit, score, selection and tail on the next context and returns a ticket;
    ``result`` synchronises that context and builds the ``TwoViewResult``.  While context A scores estimate s, the copy
    engine moves the correspondences of estimate s+1 and the host unpacks estimate s-1.  Device sampler only (the
    reference sampler is a sequential host computation); results are identical to ``two_view_arrays(...,
    sampler="device")``."""

    def __init__(self, depth: int = 2, device: Optional[int] = None):
        dev = _native.default_device() if device is None else int(device)
        self.engines = [_native.Engine(dev) for _ in range(depth)]
        self._next = 0
        self._pending = {}

    def set_score_variant(self, *a, **k):
        for e in self.engines:
            e.set_score_variant(*a, **k)

    def submit(self, camera_matrix, pts_a, pts_b, threshold, min_num_extra_inliers=0, error_aggregation_method="rms",
               max_iterations: int = 100, distance_threshold: float = 50.0, seed: int = 0, selection: str = "min_error"):
        k = self._next
        eng = self.engines[k % len(self.engines)]
        if any(t % len(self.engines) == k % len(self.engines) for t in self._pending):
            raise RuntimeError("fetch the estimate submitted to this context before submitting another one")
        self._next += 1
        n = np.asarray(pts_a).reshape(-1, 2).shape[0]
        mask, sed = _device_submit(eng, eng.pinned_out(n), camera_matrix, pts_a, pts_b, threshold, min_num_extra_inliers or 0,
                                   _agg_name(error_aggregation_method), int(max_iterations), float(distance_threshold),
                                   seed, selection)
        self._pending[k] = (eng, mask, sed, min_num_extra_inliers or 0)
        return k

    def result(self, ticket) -> TwoViewResult:
        eng, mask, sed, min_extra = self._pending.pop(ticket)
        return _device_result(eng, mask, sed, min_extra, copy=True)

    def close(self):
        for e in self.engines:
            e.close()
