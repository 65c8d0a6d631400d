"""TEST INFRASTRUCTURE ONLY — restatement of the stage right before the hot path:
the brute-force matcher (lib/feature_matching/matching.py:36-118) and its patch scores
(lib/feature_matching/ncc.py:7-54, ssd.py:7-36, util.py:8-27).

Parity status: PINNED against the unmodified reference through tests/golden/matching_known_answer.json
(tests/golden/make_golden.py) and, in the build container, live (tests/test_oracle.py).
Never imported by the product path.
"""
from __future__ import annotations

import heapq

import numpy as np


# --------------------------------------------------------------------------------------
# patch scores
# --------------------------------------------------------------------------------------
def is_within_bounds(x, y, shape, window_size):
    """util.py:8-18 (note: compares the float coordinates, the window itself uses int())."""
    h = int(window_size / 2)
    return (h <= y < shape[0] - h) and (h <= x < shape[1] - h)


def select_window(image, x, y, window_size):
    """util.py:21-27."""
    h = int(window_size / 2)
    return image[int(y) - h:int(y) + h + 1, int(x) - h:int(x) + h + 1]


def ncc_score(image_a, image_b, fa, fb, window_size=3):
    """ncc.py:7-54: 1 - NCC in [0, 2]; 2.0 when a window leaves the image or has no texture."""
    if image_a.shape != image_b.shape:
        raise ValueError("the images must have the same shape")
    if not is_within_bounds(fa[0], fa[1], image_a.shape, window_size) or not is_within_bounds(
            fb[0], fb[1], image_a.shape, window_size):
        return 2.0
    wa = select_window(image_a, fa[0], fa[1], window_size)
    wb = select_window(image_b, fb[0], fb[1], window_size)
    wa = wa - np.mean(wa)  # :36-40
    wb = wb - np.mean(wb)
    num = np.dot(wa.flatten(), wb.flatten())  # :42
    den = np.sqrt(np.sum(np.square(wa)) * np.sum(np.square(wb)))  # :43-45
    if den == 0:
        return 2.0
    nnc = num / den
    nnc *= -1.0  # :51-52
    nnc += 1.0
    return float(nnc)


def ssd_score(image_a, image_b, fa, fb, window_size=5):
    """ssd.py:7-36: mean squared difference; inf when a window leaves the image."""
    if image_a.shape != image_b.shape:
        raise ValueError("the images must have the same shape")
    if not is_within_bounds(fa[0], fa[1], image_a.shape, window_size) or not is_within_bounds(
            fb[0], fb[1], image_a.shape, window_size):
        return float("inf")
    diff = select_window(image_a, fa[0], fa[1], window_size) - select_window(image_b, fb[0], fb[1], window_size)
    sq = np.square(diff)
    return float(np.sum(sq) / sq.size)


def score_matrix(image_a, image_b, feats_a, feats_b, kind="ncc", window_size=None):
    fn = ncc_score if kind == "ncc" else ssd_score
    if window_size is None:
        window_size = 3 if kind == "ncc" else 5
    S = np.empty((len(feats_a), len(feats_b)), dtype=np.float64)
    for i, fa in enumerate(feats_a):
        for j, fb in enumerate(feats_b):
            S[i, j] = fn(image_a, image_b, fa, fb, window_size)
    return S


# --------------------------------------------------------------------------------------
# matcher
# --------------------------------------------------------------------------------------
class _M:
    """Match with the reference's ordering: by score only, strict (matching.py:22-24)."""
    __slots__ = ("a", "b", "s")

    def __init__(self, a, b, s):
        self.a, self.b, self.s = a, b, s

    def __lt__(self, other):
        return self.s < other.s


def heap_top2(scores):
    """(index of heap[0], its score, score of heap[1] or None) after pushing scores in order —
    matching.py:55-65 builds one heap per feature; :84-97 reads heap[0] and heap[1] (NOT the second
    smallest: heap[1] is the root of the left subtree)."""
    heap = []
    for b, s in enumerate(scores):
        heapq.heappush(heap, _M(0, b, s))
    if not heap:
        return -1, None, None
    return heap[0].b, heap[0].s, (heap[1].s if len(heap) > 1 else None)


def heap_top2_closed_form(scores):
    """The same three values without a heap.  heap[0] is the first minimum (strict <).  An element pushed
    at 0-based position k lands in the left subtree iff the binary form of k+1 starts with '10'; the left
    subtree then gains max(score_k, running minimum before k) (the loser of the comparison with the root),
    and heap[1] is the minimum of those."""
    s = np.asarray(scores, dtype=np.float64)
    n = len(s)
    if n == 0:
        return -1, None, None
    b0 = int(np.argmin(s))  # first occurrence
    if n == 1:
        return b0, float(s[0]), None
    run = np.minimum.accumulate(s)
    h1 = np.inf
    found = False
    for k in range(1, n):
        p = k + 1
        if (p >> (p.bit_length() - 2)) == 2:  # prefix '10'
            v = s[k] if not (s[k] < run[k - 1]) else run[k - 1]
            h1 = v if not found else min(h1, v)
            found = True
    return b0, float(s[b0]), float(h1)


def match_from_scores(S, ratio_test=False, crosscheck=False, ratio_test_threshold=0.5):
    """matching.py:36-118 given the full score matrix.  Returns [(a_index, b_index, score)] in a order."""
    S = np.asarray(S, dtype=np.float64)
    na, nb = S.shape
    per_a = []
    for a in range(na):
        b0, s0, s1 = heap_top2(S[a])
        per_a.append((a, b0, s0, s1))
    if ratio_test:  # :84-97
        kept = []
        for a, b0, s0, s1 in per_a:
            if nb > 1:
                with np.errstate(all="ignore"):
                    ok = (np.float64(s0) / np.float64(s1)) <= ratio_test_threshold
                if ok:
                    kept.append((a, b0, s0, s1))
            elif nb == 1:
                kept.append((a, b0, s0, s1))
        per_a = kept
    if crosscheck:  # :100-118
        best_for_b = {}
        for a, b0, s0, _ in per_a:
            if b0 not in best_for_b or best_for_b[b0][1] > s0:
                best_for_b[b0] = (a, s0)
        per_a = [m for m in per_a if best_for_b[m[1]][0] == m[0]]  # dataclass equality: the same match
    return [(a, b0, s0) for a, b0, s0, _ in per_a]


# --------------------------------------------------------------------------------------
# N2: Harris corners (lib/harris/harris_detector.py:11-113, lib/common/correlate.py:4-39)
# --------------------------------------------------------------------------------------
SOBEL_X = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=float)  # harris_detector.py:8


def cross_correlate(image, kernel):
    """correlate.py:4-39: per-pixel np.dot of the flattened window with the flattened kernel, zero border."""
    image, kernel = np.asarray(image), np.asarray(kernel)
    if image.ndim != 2 or kernel.ndim != 2:
        raise ValueError("Only 2D single channel images are supported")
    if kernel.shape[0] != kernel.shape[1] or (kernel.shape[0] % 2) == 0:
        raise ValueError("Only odd-sized square kernels are supported")
    ks = kernel.shape[0]
    if image.shape[0] < ks or image.shape[1] < ks:
        raise ValueError("Kernel cannot be larger than image")
    h = int(ks / 2)
    out = np.zeros(image.shape, dtype=float)
    kf = kernel.flatten()
    for r in range(h, h + image.shape[0] - ks + 1):
        for c in range(h, h + image.shape[1] - ks + 1):
            out[r, c] = np.dot(image[r - h:r + h + 1, c - h:c + h + 1].flatten(), kf)
    return out


def cornerness_image(image, block_size=2, k=0.04):
    """harris_detector.py:58-86 (before the clamp).  np.linalg.det on the stacked 2x2 matrices runs the same
    LAPACK routine per matrix as the reference's per-pixel call."""
    sy = cross_correlate(image, SOBEL_X.transpose())
    sx = cross_correlate(image, SOBEL_X)
    ix2, iy2, ixy = sx ** 2, sy ** 2, sx * sy
    height, width = image.shape
    out = np.zeros((height - int(np.around(block_size / 2)), width - int(np.around(block_size / 2))), dtype=float)
    for r in range(height - block_size):
        for c in range(width - block_size):
            a = np.sum(ix2[r:r + block_size, c:c + block_size])
            b = np.sum(ixy[r:r + block_size, c:c + block_size])
            d = np.sum(iy2[r:r + block_size, c:c + block_size])
            M = np.array([[a, b], [b, d]])
            out[r, c] = np.linalg.det(M) - k * (np.trace(M) ** 2)
    return out


def non_max_suppress(image):
    """harris_detector.py:95-104: in place, row-major — later pixels see the already suppressed values."""
    for r in range(image.shape[0]):
        for c in range(image.shape[1]):
            window = image[max(0, r - 1):min(image.shape[0], r + 2), max(0, c - 1):min(image.shape[1], c + 2)]
            if image[r, c] < np.amax(window):
                image[r, c] = 0.0


def non_max_suppress_fixed_point(image):
    """The same result as a fixed point of parallel sweeps (what the CUDA kernel iterates): a pixel survives iff
    no later neighbour is larger and every earlier neighbour is either not larger or itself suppressed."""
    v = np.asarray(image, dtype=float)
    rows, cols = v.shape
    pad = np.full((rows + 2, cols + 2), -np.inf)
    pad[1:-1, 1:-1] = v
    alive = np.ones((rows + 2, cols + 2), dtype=bool)
    sweeps = 0
    while True:
        new = np.ones((rows, cols), dtype=bool)
        for dr in (-1, 0, 1):
            for dc in (-1, 0, 1):
                if dr == 0 and dc == 0:
                    continue
                nb = pad[1 + dr:1 + dr + rows, 1 + dc:1 + dc + cols]
                larger = v < nb
                if dr < 0 or (dr == 0 and dc < 0):
                    larger &= alive[1 + dr:1 + dr + rows, 1 + dc:1 + dc + cols]
                new &= ~larger
        sweeps += 1
        if np.array_equal(new, alive[1:-1, 1:-1]):
            break
        alive[1:-1, 1:-1] = new
    return np.where(new, v, 0.0), sweeps


def harris_corners(image, num_corners=50, block_size=2, k=0.04):
    """harris_detector.py:11-55.  Returns (xy [m,2], scores [m], suppressed cornerness image).  Exact ties are
    ordered by descending flat index (numpy's own order among equal keys is unspecified)."""
    if num_corners <= 0:
        raise ValueError("num_corners needs to be at least 1")
    cim = cornerness_image(np.asarray(image), block_size, k)
    cim[cim < 0] = 0.0
    non_max_suppress(cim)
    flat = cim.ravel()
    order = np.lexsort((-np.arange(flat.size), -flat))[:num_corners]
    order = order[flat[order] != 0]
    ys, xs = np.unravel_index(order, cim.shape)
    xy = np.stack([xs.astype(float) + float(block_size) / 2.0, ys.astype(float) + float(block_size) / 2.0], 1)
    return xy, flat[order], cim


def cornerness_image_vectorised(image, block_size=2, k=0.04):
    """Whole-image form for full-size checks: the same sums, ``det = a d - b b`` written out.  For integer-valued
    (uint8) images every intermediate up to the determinant is an exact integer in fp64, so this equals the exact
    value the reference approximates through LAPACK + log/exp (np.linalg.det) to ~1e-13 relative."""
    img = np.asarray(image).astype(np.float64)
    height, width = img.shape
    sx, sy = np.zeros_like(img), np.zeros_like(img)
    acc_x, acc_y = np.zeros((height - 2, width - 2)), np.zeros((height - 2, width - 2))
    for dr in range(3):  # the nine terms in np.dot order
        for dc in range(3):
            w = img[dr:dr + height - 2, dc:dc + width - 2]
            acc_x = acc_x + w * SOBEL_X[dr, dc]
            acc_y = acc_y + w * SOBEL_X[dc, dr]
    sx[1:-1, 1:-1], sy[1:-1, 1:-1] = acc_x, acc_y
    shrink = int(np.around(block_size / 2))
    out = np.zeros((height - shrink, width - shrink))
    rr, cc = max(height - block_size, 0), max(width - block_size, 0)
    a, b, d = np.zeros((rr, cc)), np.zeros((rr, cc)), np.zeros((rr, cc))
    for dr in range(block_size):
        for dc in range(block_size):
            x, y = sx[dr:dr + rr, dc:dc + cc], sy[dr:dr + rr, dc:dc + cc]
            a, b, d = a + x * x, b + x * y, d + y * y
    tr = a + d
    out[:rr, :cc] = (a * d - b * b) - k * (tr * tr)
    return out


def harris_corners_vectorised(image, num_corners=50, block_size=2, k=0.04):
    """harris_corners with the vectorised cornerness and the fixed-point suppression (identical results to the
    sequential scan, see tests/test_front_end_oracle.py)."""
    cim = cornerness_image_vectorised(image, block_size, k)
    cim[cim < 0] = 0.0
    cim, sweeps = non_max_suppress_fixed_point(cim)
    flat = cim.ravel()
    nz = np.flatnonzero(flat)
    order = nz[np.lexsort((-nz, -flat[nz]))][:num_corners]
    ys, xs = np.unravel_index(order, cim.shape)
    xy = np.stack([xs.astype(float) + float(block_size) / 2.0, ys.astype(float) + float(block_size) / 2.0], 1)
    return xy, flat[order], cim, sweeps


# --------------------------------------------------------------------------------------
# N3: the whole of apps/sfm.py:34-186 without GUI / hydra / dataset (restated with the pieces above)
# --------------------------------------------------------------------------------------
def sfm_pipeline(image_1, image_2, K, num_harris_corners=600, ncc_window_size=9, ratio_test_threshold=0.7,
                 match_score_threshold=0.3, sed_inlier_threshold=1.5e-6, min_num_extra_inliers=10, max_iterations=2000):
    """apps/sfm.py: Harris corners (:64-71) -> brute-force NCC matching with ratio test + cross-check (:73-87) ->
    score filter (:107, :280-296) -> RANSAC essential matrix (:110-119) -> pose (:133-138) -> triangulation of the
    pairs passing the cheirality vote (:165-186).  Uses the global ``random`` state like the reference."""
    from oracle import restatement as o

    c1, _, _, _ = harris_corners_vectorised(image_1, num_harris_corners)
    c2, _, _, _ = harris_corners_vectorised(image_2, num_harris_corners)
    S = score_matrix(image_1, image_2, c1, c2, "ncc", ncc_window_size)
    matches = match_from_scores(S, True, True, ratio_test_threshold)
    matches = [m for m in matches if not (m[2] > match_score_threshold)]
    ia, ib = np.array([m[0] for m in matches], dtype=int), np.array([m[1] for m in matches], dtype=int)
    pa, pb = c1[ia], c2[ib]
    r = o.ransac_essential(K, pa[:, 0], pa[:, 1], pb[:, 0], pb[:, 1], sed_inlier_threshold, min_num_extra_inliers, "rms",
                           max_iterations)
    inl = r["inlier_indices"]
    R, t, mask, _ = o.recover_r_t_from_e(r["E"], K, pa[inl, 0], pa[inl, 1], pb[inl, 0], pb[inl, 1])
    keep = inl[mask]
    X = o.triangulate_points(pa[keep, 0], pa[keep, 1], pb[keep, 0], pb[keep, 1], K, o.tmat(R, t))
    return dict(corners_1=c1, corners_2=c2, matches=matches, E=r["E"], best_index=r["best_index"], inlier_indices=inl,
                R=R, t=t, pose_mask=mask, points=X)
