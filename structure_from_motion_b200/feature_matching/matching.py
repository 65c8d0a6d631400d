"""Drop-in mirror of lib/feature_matching/matching.py: ``Match`` (:15-24), ``ValidationStrategy``
(:27-33) and ``match_brute_force`` (:36-118) — the stage right in front of the two-view hot path
(apps/sfm.py:73-87 builds the ``Match`` list the RANSAC entry point consumes).

The reference pushes Na*Nb ``Match`` objects through ``heapq`` and evaluates the score function
once per pair in Python.  Here the score matrix, the per-feature selection, the ratio test and the
cross-check run on the GPU (csrc/sfm_match.cuh behind ``sfm_match_brute_force``), with no CPU
fallback.  A score function is recognised when it is ``calculate_ncc`` / ``calculate_ssd`` with the
two images bound — ``functools.partial``, a ``PatchScore`` or a forwarding closure like
apps/sfm.py:266-277 — and then all pairs are scored in one launch.  Any other Python callable is
user code and can only run on the host: it is called once per pair exactly as the reference does
(matching.py:55-65) and the resulting matrix goes through the same GPU selection kernels
(``sfm_match_from_scores``).
"""
from __future__ import annotations

import dataclasses
import functools
import math
from enum import Enum
from typing import Callable, List, NewType, Set

import numpy as np

from ..common.feature import Feature
from . import _patch

# Interface definition of a matching function.
ScoreFunction = NewType("ScoreFunction", Callable[[Feature, Feature], float])


@dataclasses.dataclass
class Match:
    a_index: int = -1
    b_index: int = -1
    # A lower score indicates a better match in all cases.
    match_score: float = math.inf

    def __lt__(self, other) -> bool:
        return self.match_score < other.match_score


class ValidationStrategy(Enum):
    """Validation strategy when matching features (matching.py:27-33)."""

    CROSSCHECK = 1
    RATIO_TEST = 2


class PatchScore:
    """A score function with its images attached: ``PatchScore(image_a, image_b, "ncc", window_size=9)``.
    Callable on one feature pair like any ``ScoreFunction``; ``match_brute_force`` unpacks it."""

    def __init__(self, image_a, image_b, kind: str = "ncc", window_size: int | None = None):
        if kind not in ("ncc", "ssd"):
            raise ValueError("kind must be 'ncc' or 'ssd'")
        self.image_a, self.image_b = _patch.check_images(image_a, image_b)
        self.kind = kind
        self.window_size = int(window_size) if window_size is not None else (3 if kind == "ncc" else 5)

    def __call__(self, feature_a: Feature, feature_b: Feature) -> float:
        return _patch.single_score(self.kind, self.image_a, self.image_b, feature_a, feature_b, self.window_size)


def _kind_and_window(fn):
    """(kind, bound positional args, window_size or None) if fn is calculate_ncc / calculate_ssd, possibly
    through functools.partial; else None."""
    from . import ncc, ssd

    args, kw = (), {}
    if isinstance(fn, functools.partial):
        args, kw, fn = fn.args, fn.keywords, fn.func
    if fn is ncc.calculate_ncc:
        kind = "ncc"
    elif fn is ssd.calculate_ssd:
        kind = "ssd"
    else:
        return None
    if set(kw) - {"image_a", "image_b", "window_size"}:
        return None
    return kind, args, kw


def _recognise(score_function):
    """PatchScore for the score functions the GPU can evaluate itself, else None."""
    if isinstance(score_function, PatchScore):
        return score_function
    try:
        direct = _kind_and_window(score_function)
        if direct is not None:  # partial(calculate_ncc, image_a, image_b, window_size=w)
            kind, args, kw = direct
            images = list(args) + [kw[k] for k in ("image_a", "image_b") if k in kw]
            if len(images) != 2 or len(args) > 2:
                return None
            return PatchScore(images[0], images[1], kind, kw.get("window_size"))
        # a closure that forwards to calculate_ncc / calculate_ssd with two captured images
        # (apps/sfm.py:266-277: ``full_score_function(image_a, image_b, feature_a, feature_b)``)
        cells = getattr(score_function, "__closure__", None)
        code = getattr(score_function, "__code__", None)
        if not cells or code is None or code.co_argcount != 2:
            return None
        free = dict(zip(code.co_freevars, (c.cell_contents for c in cells)))
        images = [(n, v) for n, v in free.items() if isinstance(v, np.ndarray)]
        inner = [v for v in free.values() if callable(v) and _kind_and_window(v) is not None]
        if len(images) != 2 or len(inner) != 1 or len(free) != 3:
            return None
        kind, args, kw = _kind_and_window(inner[0])
        if args or set(kw) - {"window_size"}:
            return None
        return PatchScore(images[0][1], images[1][1], kind, kw.get("window_size"))
    except (ValueError, TypeError):
        return None


def _agrees(candidate: PatchScore, score_function, features_a, features_b, S) -> bool:
    """A structurally recognised closure is trusted only if it reproduces the matrix on probe pairs
    (it could, for instance, pass the images in the other order)."""
    if isinstance(score_function, (PatchScore, functools.partial)):
        return True
    na, nb = S.shape
    for a, b in {(0, 0), (na - 1, nb - 1), (na // 2, nb // 3)}:
        v = float(score_function(features_a[a], features_b[b]))
        if not (v == S[a, b] or (math.isnan(v) and math.isnan(S[a, b]))):
            return False
    return True


def match_brute_force(
    features_a: List[Feature],
    features_b: List[Feature],
    score_function: ScoreFunction,
    *,
    validation_strategies: ValidationStrategy | Set[ValidationStrategy] | None = None,
    ratio_test_threshold: float = 0.5,
) -> List[Match]:
    """Match two lists of features pairwise (matching.py:36-81).

    Returns, in the order of ``features_a``, the best match (lowest score, the first one on ties) of every
    feature of A that survives the validations: RATIO_TEST keeps a feature iff ``heap[0]/heap[1] <=
    ratio_test_threshold`` where heap is the reference's per-feature ``heapq`` (matching.py:84-97 — heap[1]
    is the root of the left subtree, not necessarily the second-best score); CROSSCHECK keeps a match iff it
    is the lowest-scored (earliest on ties) surviving match of its B feature (matching.py:100-118)."""
    from .. import _native

    if validation_strategies is None:
        validation_strategies = set()
    elif not isinstance(validation_strategies, set):
        validation_strategies = set([validation_strategies])
    ratio = ValidationStrategy.RATIO_TEST in validation_strategies
    cross = ValidationStrategy.CROSSCHECK in validation_strategies
    na, nb = len(features_a), len(features_b)
    if na == 0:
        return []
    if nb == 0:
        if ratio:
            return []  # matching.py:89-96 drops features with an empty heap
        raise IndexError("list index out of range")  # matching.py:79 / :105 index an empty heap

    eng = _native.get_engine()
    best_b = best_s = keep = None
    patch = _recognise(score_function)
    if patch is not None:
        fa, fb = _patch.feature_array(features_a), _patch.feature_array(features_b)
        # a forwarding closure is only trusted after probing a few pairs against the score matrix (see _agrees)
        probe = not isinstance(score_function, (PatchScore, functools.partial))
        best_b, best_s, keep, S = eng.match_brute_force(
            patch.image_a, patch.image_b, fa, fb, kind=patch.kind, window=patch.window_size, ratio_test=ratio,
            crosscheck=cross, ratio_threshold=ratio_test_threshold, want_scores=probe)
        if probe and not _agrees(patch, score_function, features_a, features_b, S):
            best_b = None
    if best_b is None:
        S = np.empty((na, nb), dtype=np.float64)
        for a_index, feature_a in enumerate(features_a):  # matching.py:55-65, user code on the host
            for b_index, feature_b in enumerate(features_b):
                S[a_index, b_index] = score_function(feature_a, feature_b)
        best_b, best_s, keep = eng.match_from_scores(S, ratio_test=ratio, crosscheck=cross,
                                                     ratio_threshold=ratio_test_threshold)
    return [Match(a_index=int(a), b_index=int(best_b[a]), match_score=float(best_s[a])) for a in np.flatnonzero(keep)]
