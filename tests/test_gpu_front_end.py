"""GPU parity of the stages in front of the hot path: brute-force matcher with NCC / SSD
patch scores, Harris corners, cross-correlation — against the golden vectors of the unmodified reference and the
oracle (oracle/front_end.py), through the C ABI and through the reference's own call signatures."""
import functools
import json
import os

import numpy as np
import pytest

from oracle import front_end as fe
from structure_from_motion_b200 import _native

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# NCC: the reference's np.dot goes through OpenBLAS ddot, whose summation order over the w*w window terms is not
# the kernel's sequential FMA chain; everything else in the score is the same IEEE operations.  |score| <= 2.
NCC_ATOL = 1e-13


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def mg():
    d = load("matching_known_answer.json")
    d["image_a"] = np.array(d["image_a"], dtype=np.uint8)
    d["image_b"] = np.array(d["image_b"], dtype=np.uint8)
    d["feats_a"], d["feats_b"] = np.array(d["feats_a"]), np.array(d["feats_b"])
    return d


def _features(xy):
    from lib.common.feature import Feature

    return [Feature(x=float(x), y=float(y)) for x, y in xy]


def test_score_matrix_against_reference(engine, mg):
    for case in mg["cases"]:
        if case["scores"] is None:
            continue
        ref = np.array(case["scores"])
        _, _, _, S = engine.match_brute_force(mg["image_a"], mg["image_b"], mg["feats_a"], mg["feats_b"], kind=case["kind"],
                                              window=case["window"], want_scores=True)
        if case["kind"] == "ssd":
            assert np.array_equal(S, ref)  # uint8 wrap-around arithmetic and +inf outside: bit-exact
        else:
            assert np.array_equal(S == 2.0, ref == 2.0)
            assert np.abs(S - ref).max() <= NCC_ATOL
        # float64 images: SSD without the uint8 wrap-around
        if case["kind"] == "ssd":
            fa, fb = mg["image_a"].astype(np.float64), mg["image_b"].astype(np.float64)
            _, _, _, Sf = engine.match_brute_force(fa, fb, mg["feats_a"], mg["feats_b"], kind="ssd", window=case["window"],
                                                   want_scores=True)
            ref_f = fe.score_matrix(fa, fb, mg["feats_a"], mg["feats_b"], "ssd", case["window"])
            assert np.array_equal(np.isinf(Sf), np.isinf(ref_f))
            ok = np.isfinite(ref_f)
            assert np.allclose(Sf[ok], ref_f[ok], rtol=1e-14, atol=0)  # integer-valued: exact up to the final division


def test_even_window_sizes_follow_the_reference_slicing(engine, mg):
    """util.py:21-27 cuts center +- int(w/2): w = 4 is a 5x5 patch, w = 1 a single pixel."""
    for kind, w in [("ncc", 4), ("ssd", 2), ("ssd", 1), ("ncc", 8)]:
        ref = fe.score_matrix(mg["image_a"], mg["image_b"], mg["feats_a"][:20], mg["feats_b"][:19], kind, w)
        _, _, _, S = engine.match_brute_force(mg["image_a"], mg["image_b"], mg["feats_a"][:20], mg["feats_b"][:19], kind=kind,
                                              window=w, want_scores=True)
        assert np.array_equal(np.isfinite(S), np.isfinite(ref))
        ok = np.isfinite(ref)
        assert np.abs(S[ok] - ref[ok]).max() <= (NCC_ATOL if kind == "ncc" else 0.0)


def test_matches_against_reference_all_validation_modes(engine, mg):
    for case in mg["cases"]:
        ratio, cross = "RATIO_TEST" in case["strategies"], "CROSSCHECK" in case["strategies"]
        bb, bs, keep, _ = engine.match_brute_force(mg["image_a"], mg["image_b"], mg["feats_a"], mg["feats_b"],
                                                   kind=case["kind"], window=case["window"], ratio_test=ratio, crosscheck=cross,
                                                   ratio_threshold=case["ratio"])
        got = [(int(a), int(bb[a])) for a in np.flatnonzero(keep)]
        assert got == [(m[0], m[1]) for m in case["matches"]], (case["kind"], case["strategies"])
        ref_s = np.array([m[2] for m in case["matches"]])
        assert np.allclose(bs[keep], ref_s, rtol=0, atol=NCC_ATOL if case["kind"] == "ncc" else 0)


def test_selection_is_exact_given_the_scores(engine, mg):
    """heap[0], heap[1], ratio test and cross-check on the reference's own score matrices: identical matches."""
    scores = {}
    for case in mg["cases"]:
        key = (case["kind"], case["window"])
        if case["scores"] is not None:
            scores[key] = np.array(case["scores"])
        bb, bs, keep = engine.match_from_scores(scores[key], "RATIO_TEST" in case["strategies"],
                                                "CROSSCHECK" in case["strategies"], case["ratio"])
        got = [[int(a), int(bb[a]), float(bs[a])] for a in np.flatnonzero(keep)]
        assert got == case["matches"]


@pytest.mark.parametrize("na,nb", [(1, 1), (3, 1), (1, 2), (5, 3), (64, 33), (300, 257)])
def test_selection_random_with_ties(engine, na, nb):
    rng = np.random.default_rng(na * 1000 + nb)
    for mode in range(3):
        if mode == 0:
            S = rng.integers(0, 5, (na, nb)).astype(float)  # heavy ties, zeros (0/0 ratio = NaN fails)
        elif mode == 1:
            S = rng.random((na, nb)) * 2
            S[rng.random((na, nb)) < 0.3] = 2.0
        else:
            S = rng.random((na, nb))
            S[rng.random((na, nb)) < 0.3] = np.inf
            S[0] = np.inf
        for ratio, cross in [(False, False), (True, False), (False, True), (True, True)]:
            want = fe.match_from_scores(S, ratio, cross, 0.8)
            bb, bs, keep = engine.match_from_scores(S, ratio, cross, 0.8)
            got = [(int(a), int(bb[a]), float(bs[a])) for a in np.flatnonzero(keep)]
            assert got == want, (mode, ratio, cross)


def test_match_brute_force_through_the_reference_signature(mg):
    """functools.partial, the forwarding closure of apps/sfm.py:266-277 and an arbitrary Python callable."""
    from lib.feature_matching import matching, ncc, ssd

    fa, fb = _features(mg["feats_a"]), _features(mg["feats_b"])
    img_a, img_b = mg["image_a"], mg["image_b"]
    VS = matching.ValidationStrategy

    def create_score_function(image_a, image_b, full_score_function):  # apps/sfm.py:266-277
        def ssd_score(feature_a, feature_b):
            return full_score_function(image_a, image_b, feature_a, feature_b)

        return ssd_score

    for case in mg["cases"]:
        fn = ncc.calculate_ncc if case["kind"] == "ncc" else ssd.calculate_ssd
        strategies = {VS[s] for s in case["strategies"]} or None
        ref_pairs = [(m[0], m[1]) for m in case["matches"]]
        for score in (functools.partial(fn, img_a, img_b, window_size=case["window"]),
                      create_score_function(img_a, img_b, functools.partial(fn, window_size=case["window"])),
                      matching.PatchScore(img_a, img_b, case["kind"], case["window"])):
            assert matching._recognise(score) is not None
            got = matching.match_brute_force(fa, fb, score, validation_strategies=strategies,
                                             ratio_test_threshold=case["ratio"])
            assert all(isinstance(m, matching.Match) for m in got)
            assert [(m.a_index, m.b_index) for m in got] == ref_pairs
    # an arbitrary callable: scored on the host by the user's code, selected on the GPU
    case = mg["cases"][4]
    S = np.array(mg["cases"][0]["scores"])
    index_a = {id(f): i for i, f in enumerate(fa)}
    index_b = {id(f): i for i, f in enumerate(fb)}
    got = matching.match_brute_force(fa, fb, lambda a, b: S[index_a[id(a)], index_b[id(b)]],
                                     validation_strategies={VS.RATIO_TEST, VS.CROSSCHECK}, ratio_test_threshold=case["ratio"])
    assert [[m.a_index, m.b_index, m.match_score] for m in got] == case["matches"]
    # a closure that swaps the images is not trusted: it goes down the host-callable path and still gives ITS answer
    swapped = create_score_function(img_b, img_a, functools.partial(ncc.calculate_ncc, window_size=9))
    got = matching.match_brute_force(fa[:6], fb[:5], lambda a, b: swapped(a, b))
    want = fe.match_from_scores(fe.score_matrix(img_b, img_a, mg["feats_a"][:6], mg["feats_b"][:5], "ncc", 9))
    assert [(m.a_index, m.b_index) for m in got] == [(a, b) for a, b, _ in want]


def test_matcher_edge_cases(mg):
    from lib.feature_matching import matching, ncc, ssd

    fa, fb = _features(mg["feats_a"]), _features(mg["feats_b"])
    img_a, img_b = mg["image_a"], mg["image_b"]
    score = functools.partial(ncc.calculate_ncc, img_a, img_b)
    VS = matching.ValidationStrategy
    assert matching.match_brute_force([], fb, score) == []
    with pytest.raises(IndexError):  # matching.py:79 indexes an empty heap
        matching.match_brute_force(fa, [], score)
    assert matching.match_brute_force(fa, [], score, validation_strategies=VS.RATIO_TEST) == []
    one = matching.match_brute_force(fa, fb[:1], score, validation_strategies=VS.RATIO_TEST)  # matching.py:95-96
    assert [m.b_index for m in one] == [0] * len(fa)
    # single-pair calls of the score functions themselves
    S_ncc = np.array(mg["cases"][5]["scores"])   # ncc, window 3 (the default)
    S_ssd = np.array(mg["cases"][10]["scores"])  # ssd, window 5 (the default)
    for a, b in [(0, 0), (3, 4), (5, 7), (12, 1), (20, 20)]:
        assert abs(ncc.calculate_ncc(img_a, img_b, fa[a], fb[b]) - S_ncc[a, b]) <= NCC_ATOL
        assert ssd.calculate_ssd(img_a, img_b, fa[a], fb[b]) == S_ssd[a, b]
    with pytest.raises(ValueError, match="same shape"):
        ncc.calculate_ncc(img_a, img_b[:-1], fa[0], fb[0])
    with pytest.raises(ValueError, match="same shape"):
        ssd.calculate_ssd(img_a[:, :-1], img_b, fa[0], fb[0])
    flat = np.full((20, 20), 7, dtype=np.uint8)  # no texture: denominator 0 -> 2.0 (ncc.py:47-48)
    from lib.common.feature import Feature

    assert ncc.calculate_ncc(flat, flat, Feature(10.0, 10.0), Feature(9.0, 9.0)) == 2.0
    assert ssd.calculate_ssd(flat, flat, Feature(10.0, 10.0), Feature(9.0, 9.0)) == 0.0


def test_matcher_full_size_properties(engine):
    """apps/config/config.yaml sizes and beyond: 2000 x 2000 features, 9x9 NCC windows, both validations."""
    rng = np.random.default_rng(7)
    h, w = 480, 640
    base = rng.integers(0, 256, (h + 8, w + 8)).astype(np.float64)
    sm = sum(base[i:i + h, j:j + w] for i in range(3) for j in range(3)) / 9.0
    img_a = np.clip(np.round(sm), 0, 255).astype(np.uint8)
    img_b = np.clip(np.round(np.roll(sm, (2, 3), (0, 1)) + rng.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    n = 2000
    fa = np.stack([rng.integers(0, w, n), rng.integers(0, h, n)], 1).astype(np.float64)
    fb = np.stack([np.clip(fa[:, 0] + 3, 0, w - 1), np.clip(fa[:, 1] + 2, 0, h - 1)], 1)[rng.permutation(n)]
    bb, bs, keep, S = engine.match_brute_force(img_a, img_b, fa, fb, kind="ncc", window=9, ratio_test=True, crosscheck=True,
                                               ratio_threshold=0.7, want_scores=True)
    # sampled entries against the oracle's per-pair score
    for a, b in zip(rng.integers(0, n, 200), rng.integers(0, n, 200)):
        assert abs(S[a, b] - fe.ncc_score(img_a, img_b, fa[a], fb[b], 9)) <= NCC_ATOL
    # the selection is exact given the matrix
    assert np.array_equal(bb, np.argmin(S, axis=1)) and np.array_equal(bs, S[np.arange(n), bb])
    want = fe.match_from_scores(S[:300], True, False, 0.7)  # the heap-based restatement on a slice (ratio test is per row)
    bb2, bs2, keep2 = engine.match_from_scores(S[:300], True, False, 0.7)
    assert [(int(a), int(bb2[a]), float(bs2[a])) for a in np.flatnonzero(keep2)] == want
    # cross-check: an injection; every kept match is the best of its B feature among the ratio-test survivors
    kept_b = bb[keep]
    assert len(np.unique(kept_b)) == len(kept_b) and keep.sum() > n // 4
    _, _, keep_ratio, _ = engine.match_brute_force(img_a, img_b, fa, fb, kind="ncc", window=9, ratio_test=True,
                                                   ratio_threshold=0.7)
    assert (keep <= keep_ratio).all()
    for a in np.flatnonzero(keep)[:200]:
        rivals = np.flatnonzero(keep_ratio & (bb == bb[a]))
        assert bs[a] == bs[rivals].min() and a == rivals[bs[rivals] == bs[a]].min()
    # most of the planted correspondences are recovered
    planted = {(int(x), int(y)) for x, y in fb}
    hits = sum((int(np.clip(fa[a, 0] + 3, 0, w - 1)), int(np.clip(fa[a, 1] + 2, 0, h - 1))) == (int(fb[bb[a], 0]), int(fb[bb[a], 1]))
               for a in np.flatnonzero(keep))
    assert hits >= 0.9 * keep.sum() and planted


# ------------------------------------------------------------------------------------------------
# Harris
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hg():
    d = load("harris_known_answer.json")
    d["image"] = np.array(d["image"], dtype=np.uint8)
    return d


def _same_corners_up_to_near_ties(xy, ref_xy, ref_score, rtol):
    """Identical sequences, except that corners whose reference cornerness values agree within rtol may be permuted."""
    assert len(xy) == len(ref_xy)
    i = 0
    while i < len(ref_xy):
        j = i + 1
        while j < len(ref_xy) and abs(ref_score[j] - ref_score[j - 1]) <= rtol * abs(ref_score[i]):
            j += 1
        assert sorted(map(tuple, xy[i:j])) == sorted(map(tuple, ref_xy[i:j])), (i, j)
        i = j


def test_harris_against_reference(engine, hg):
    img = hg["image"]
    for case in hg["cases"]:
        ref_cim = np.array(case["cornerness"])
        xy, score, extra = engine.harris_corners(img, case["num_corners"], case["block_size"], case["k"], want_cornerness=True)
        cim = extra["cornerness"]
        assert cim.shape == ref_cim.shape
        # the reference's determinant goes through LAPACK + log/exp: 1e-13 of the terms that cancel
        scale = np.abs(np.array(case["cornerness_raw"])).max()
        assert np.array_equal(cim != 0, ref_cim != 0)  # suppression pattern identical
        assert np.abs(cim - ref_cim).max() <= 1e-13 * scale
        ref_xy = np.array(case["corners"]).reshape(-1, 2)
        off = case["block_size"] / 2.0
        ref_score = np.array([ref_cim[int(y - off), int(x - off)] for x, y in ref_xy])
        if case["num_corners"] < 1000:  # the cut-off must not fall inside a near-tie for the strict comparison
            _same_corners_up_to_near_ties(xy, ref_xy, ref_score, 1e-12)
        else:
            assert sorted(map(tuple, xy)) == sorted(map(tuple, ref_xy))
        assert np.all(np.diff(score) <= 0) and np.all(score > 0)
        # bit-exact against the exact-integer restatement, order included
        xy_v, score_v, cim_v, _ = fe.harris_corners_vectorised(img, case["num_corners"], case["block_size"], case["k"])
        assert np.array_equal(cim, cim_v) and np.array_equal(xy, xy_v) and np.array_equal(score, score_v)


def test_harris_reference_fixtures_through_the_reference_signature(hg):
    from lib.common.feature import Feature
    from lib.harris import harris_detector as harris

    sq = np.array(hg["square"]["image"], dtype=np.uint8)
    corners = harris.detect_harris_corners(sq, num_corners=4)
    assert all(isinstance(c, Feature) for c in corners)
    assert sorted((c.x, c.y) for c in corners) == sorted(map(tuple, hg["square"]["corners"]))
    r = hg["rectangle"]  # test_harris_detector.py:14-32 (a float image)
    img = np.zeros(r["shape"])
    img[r["fill"][0]:r["fill"][1], r["fill"][2]:r["fill"][3]] = 255.0
    corners = harris.detect_harris_corners(img)
    assert sorted((c.x, c.y) for c in corners) == sorted(map(tuple, r["corners"]))
    for (ey, ex), c in zip(sorted(r["expected_yx"]), sorted((c.y, c.x) for c in corners)):
        assert np.allclose((ey, ex), c, atol=1.0)
    with pytest.raises(ValueError, match="at least 1"):
        harris.detect_harris_corners(sq, num_corners=0)
    with pytest.raises(ValueError):
        harris.detect_harris_corners(np.zeros((4, 4, 3)))
    assert harris.detect_harris_corners(np.zeros((12, 12), dtype=np.uint8)) == []  # zero cornerness is never returned


def test_harris_full_size_bit_exact_against_vectorised_oracle(engine):
    """640 x 480 (the Middlebury temple size), 600 corners as in apps/config/config.yaml."""
    rng = np.random.default_rng(3)
    h, w = 480, 640
    base = rng.integers(0, 256, (h + 8, w + 8)).astype(np.float64)
    sm = sum(base[i:i + h, j:j + w] for i in range(5) for j in range(5)) / 25.0
    img = np.clip(np.round((sm - sm.min()) * 255 / (sm.max() - sm.min())), 0, 255).astype(np.uint8)
    img[100:200, 150:300] = 250
    img[300:400, 350:500] = 5
    for bs, num in [(2, 600), (3, 50), (5, 100000)]:
        xy, score, extra = engine.harris_corners(img, num, bs, 0.04, want_cornerness=True)
        xy_v, score_v, cim_v, sweeps = fe.harris_corners_vectorised(img, num, bs, 0.04)
        assert np.array_equal(extra["cornerness"], cim_v)
        assert np.array_equal(xy, xy_v) and np.array_equal(score, score_v)
        assert extra["nms_sweeps"] >= sweeps - 1


def test_harris_suppression_with_long_dependency_chains(engine):
    """Plateaus and ramps: the in-place scan of harris_detector.py:95-104 alternates along a ramp."""
    img = np.zeros((40, 64), dtype=np.uint8)
    img[:, :] = (np.arange(64)[None, :] * 3 + np.arange(40)[:, None] * 2).astype(np.uint8)
    img[10:30, 20:44] = 200
    xy, score, extra = engine.harris_corners(img, 500, 2, 0.04, want_cornerness=True)
    raw = fe.cornerness_image_vectorised(img, 2, 0.04)
    raw[raw < 0] = 0.0
    seq = raw.copy()
    fe.non_max_suppress(seq)  # the reference's sequential scan
    assert np.array_equal(extra["cornerness"], seq)


def test_cross_correlate(engine, hg):
    from lib.blur import gaussian
    from lib.common import correlate

    d, img = hg["correlate"], hg["image"]
    kern = np.array(d["kernel"])
    assert np.array_equal(gaussian.create_gaussian_kernel(5, 1.2), kern)
    assert np.array_equal(correlate.cross_correlate(img[:20, :24], fe.SOBEL_X), np.array(d["sobel_u8"]))  # integers: exact
    out = correlate.cross_correlate(img[:20, :24].astype(np.float64) / 255.0, kern)
    assert np.allclose(out, np.array(d["result"]), rtol=1e-14, atol=1e-16)  # np.dot's summation order
    assert np.all(out[:2] == 0) and np.all(out[:, -2:] == 0)  # zero "same" border
    with pytest.raises(ValueError, match="odd-sized"):
        correlate.cross_correlate(img, np.ones((4, 4)))
    with pytest.raises(ValueError, match="larger than image"):
        correlate.cross_correlate(img[:2, :2], np.ones((3, 3)))
    with pytest.raises(ValueError, match="2D"):
        correlate.cross_correlate(np.zeros((3, 3, 3)), np.ones((3, 3)))
    with pytest.raises(ValueError):
        gaussian.create_gaussian_kernel(4, 1.0)


# ------------------------------------------------------------------------------------------------
# N3: the whole pipeline of apps/sfm.py on a rendered image pair
# ------------------------------------------------------------------------------------------------
def test_headless_sfm_app_against_oracle_pipeline():
    import random

    from apps import sfm as app
    from structure_from_motion_b200.scenes import make_image_pair

    img1, img2, K, R_true, t_true = make_image_pair(0)
    cfg = app.load_config(num_harris_corners=400, sed_inlier_threshold=1e-5, min_num_extra_inliers=60, max_iterations=1000)
    try:
        import cv2  # noqa: F401

        check = True
    except ImportError:
        check = False
    random.seed(5)
    res = app.run_sfm(img1, img2, K, cfg, check_opencv=check)  # raises if OpenCV's pose / points disagree (apps/sfm.py:140-202)
    state_after = random.getstate()
    random.seed(5)
    ref = fe.sfm_pipeline(img1, img2, K, num_harris_corners=400, sed_inlier_threshold=1e-5, min_num_extra_inliers=60,
                          max_iterations=1000)
    assert random.getstate() == state_after  # the global RNG advanced exactly as in the reference
    assert np.array_equal(np.array([[c.x, c.y] for c in res.corners_1]), ref["corners_1"])
    assert np.array_equal(np.array([[c.x, c.y] for c in res.corners_2]), ref["corners_2"])
    assert [(m.a_index, m.b_index) for m in res.matches] == [(a, b) for a, b, _ in ref["matches"]]
    assert np.allclose([m.match_score for m in res.matches], [s for _, _, s in ref["matches"]], rtol=0, atol=NCC_ATOL)
    assert np.allclose(res.e, ref["E"], rtol=1e-6, atol=1e-9)
    pa, pb = ref["corners_1"], ref["corners_2"]
    want_pairs = [(tuple(pa[ref["matches"][i][0]]), tuple(pb[ref["matches"][i][1]])) for i in ref["inlier_indices"]]
    assert [((p[0].x, p[0].y), (p[1].x, p[1].y)) for p in res.inlier_feature_pairs] == want_pairs
    assert np.allclose(res.r, ref["R"], atol=1e-6) and np.allclose(res.t, ref["t"], atol=1e-6)
    assert np.array_equal(res.inlier_mask, ref["pose_mask"])
    assert np.allclose(res.world_points, ref["points"], rtol=1e-6, atol=1e-9)
    # and the answer is right: rotation within 0.5 deg, translation direction within 3 deg of the rendered truth
    rot_err = np.degrees(np.arccos(np.clip((np.trace(res.r.T @ R_true) - 1) / 2, -1, 1)))
    t_err = np.degrees(np.arccos(np.clip(res.t @ t_true / np.linalg.norm(t_true) / np.linalg.norm(res.t), -1, 1)))
    assert rot_err < 0.5 and t_err < 3.0, (rot_err, t_err)
    assert len(res.inlier_feature_pairs) >= 68 and (res.world_points[:, 2] > 0).all()


def test_array_native_image_pair_pipeline_equals_the_list_based_app(engine):
    """two_view.image_pair_arrays (no Feature / Match objects) against apps/sfm.run_sfm on the same image pair."""
    import random

    from apps import sfm as app
    from structure_from_motion_b200 import two_view
    from structure_from_motion_b200.scenes import make_image_pair

    img1, img2, K, *_ = make_image_pair(4, h=240, w=320)
    kw = dict(num_harris_corners=300, sed_inlier_threshold=2e-5, min_num_extra_inliers=30, max_iterations=500)
    random.seed(11)
    res = app.run_sfm(img1, img2, K, app.load_config(**kw))
    state = random.getstate()
    random.seed(11)
    arr = two_view.image_pair_arrays(img1, img2, K, engine=engine, **kw)
    assert random.getstate() == state
    assert np.array_equal(arr.corners_a, np.array([[c.x, c.y] for c in res.corners_1]))
    assert [(int(a), int(b)) for a, b in zip(arr.match_a, arr.match_b)] == [(m.a_index, m.b_index) for m in res.matches]
    assert np.array_equal(arr.two_view.ransac.E, res.e)
    assert np.array_equal(arr.two_view.R, res.r) and np.array_equal(arr.two_view.t, res.t)
    # the list pipeline triangulates the pairs that pass the vote, in the reference's inlier order; the array pipeline
    # reports every inlier in ascending index order with NaN for the ones that fail: same set of points
    got = arr.two_view.points[arr.two_view.passing]
    assert len(got) == len(res.world_points)
    key = lambda X: X[np.lexsort(X.T[::-1])]  # noqa: E731
    assert np.allclose(key(got), key(res.world_points), rtol=1e-9, atol=1e-12)
