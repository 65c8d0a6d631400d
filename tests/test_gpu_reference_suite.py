"""The reference's own front-end unit tests, restated case by case and run against the B200 implementation through the
reference's import paths: lib/feature_matching/tests/{test_matching,test_ncc,test_ssd,test_util}.py,
lib/common/tests/test_correlate.py, lib/blur/tests/test_gaussian.py, lib/harris/tests/test_harris_detector.py.
(The epipolar / RANSAC tests of the reference are restated in tests/test_gpu_golden.py.)"""
from functools import partial

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_match_features_no_validation():  # test_matching.py:8-36
    from lib.common.feature import Feature
    from lib.feature_matching import matching

    features_a = [Feature(1, 1), Feature(2, 2), Feature(3, 3)]
    features_b = [Feature(4, 4), Feature(5, 5), Feature(6, 6), Feature(7, 7)]
    table = {0: {0: 10, 1: 20, 2: 30, 3: 7}, 1: {0: 30, 1: 9, 2: 20, 3: 15}, 2: {0: 20, 1: 30, 2: 8, 3: 31}}

    def score(feature_a, feature_b):
        return table[features_a.index(feature_a)][features_b.index(feature_b)]

    matches = matching.match_brute_force(features_a, features_b, score)
    assert matches == [matching.Match(a_index=0, b_index=3, match_score=7), matching.Match(a_index=1, b_index=1, match_score=9),
                       matching.Match(a_index=2, b_index=2, match_score=8)]


def test_match_features_ratio_test():  # test_matching.py:38-63
    from lib.common.feature import Feature
    from lib.feature_matching import matching

    features_a = [Feature(1, 1), Feature(2, 2)]
    features_b = [Feature(3, 3), Feature(4, 4), Feature(5, 5)]
    table = {0: {0: 10, 1: 5, 2: 20},   # ratio exactly 0.5, the threshold
             1: {0: 10, 1: 6, 2: 7}}    # ratio above 0.5

    def score(feature_a, feature_b):
        return table[features_a.index(feature_a)][features_b.index(feature_b)]

    matches = matching.match_brute_force(features_a, features_b, score,
                                         validation_strategies=matching.ValidationStrategy.RATIO_TEST, ratio_test_threshold=0.5)
    assert matches == [matching.Match(0, 1, 5)]


def test_ncc():  # test_ncc.py:10-56
    from lib.common.feature import Feature
    from lib.feature_matching import ncc

    image_a = np.array([[1, 2, 3, 4, 5], [6, 7, 8, 9, 10], [9, 8, 7, 6, 5], [4, 3, 2, 1, 0], [1, 2, 3, 4, 5]])
    np.testing.assert_allclose(0.0, ncc.calculate_ncc(image_a, np.copy(image_a), Feature(2, 2), Feature(2, 2), 5), atol=1e-10)
    np.testing.assert_allclose(2.0, ncc.calculate_ncc(image_a, -np.copy(image_a), Feature(2, 2), Feature(2, 2), 5))
    np.random.seed(55)
    for _ in range(100):
        a = np.random.rand(5, 5) + np.random.randint(-100, 100)
        b = np.random.rand(5, 5) + np.random.randint(-100, 100)
        s = ncc.calculate_ncc(a, b, Feature(2, 2), Feature(2, 2), 5)
        assert -1e-8 <= s < 2.0 + 1e-8


def test_ssd():  # test_ssd.py:10-58
    from lib.common import feature
    from lib.feature_matching import ssd

    image_a = np.zeros((6, 6), dtype=int)
    image_a[:3, :3] = np.arange(1, 10).reshape(3, 3)
    image_b = np.zeros((6, 6), dtype=int)
    image_b[3:, 3:] = np.arange(9, 0, -1).reshape(3, 3)
    calc = partial(ssd.calculate_ssd, image_a, image_b, window_size=3)
    assert calc(feature.Feature(1, 1), feature.Feature(4, 4)) == (64 + 36 + 16 + 4 + 0 + 4 + 16 + 36 + 64) / 3 / 3
    assert calc(feature.Feature(0, 0), feature.Feature(4, 4)) == np.inf  # out-of-bounds features
    assert calc(feature.Feature(1, 1), feature.Feature(5, 5)) == np.inf


def test_is_within_bounds():  # test_util.py:9-24
    from lib.common import feature
    from lib.feature_matching import util

    inside = partial(util.is_within_bounds, image_shape=(100, 200), window_size=5)
    assert inside(feature.Feature(2, 2)) and not inside(feature.Feature(1, 2)) and not inside(feature.Feature(2, 1))
    assert inside(feature.Feature(y=97, x=197)) and not inside(feature.Feature(y=98, x=197))
    assert not inside(feature.Feature(y=97, x=198))


def test_cross_correlate():  # test_correlate.py:9-42
    from lib.common import correlate

    image = np.ones((5, 10), dtype=float)
    out = correlate.cross_correlate(image, np.ones((3, 3), dtype=float))
    assert np.allclose(out[1:-1, 1:-1], 9) and np.allclose(out[0], 0) and np.allclose(out[:, -1], 0)
    out = correlate.cross_correlate(image, np.ones((5, 5), dtype=float))
    assert np.allclose(out[2:-2, 2:-2], 25) and np.allclose(out[:2], 0) and np.allclose(out[:, -2:], 0)
    image = np.array([[1, 5, 4, 3, 7], [2, 5, 7, 4, -10], [9, -5, 4, 3, 2]], dtype=float)
    kernel = np.array([[1, -2, 3], [2, 1, 0], [7, -5, 1]], dtype=float)
    out = correlate.cross_correlate(image, kernel)
    assert out[1, 1] == 1 * 1 + 5 * -2 + 4 * 3 + 2 * 2 + 5 * 1 + 7 * 0 + 9 * 7 + -5 * -5 + 4 * 1
    assert out[1, 2] == 5 * 1 + 4 * -2 + 3 * 3 + 5 * 2 + 7 * 1 + 4 * 0 + -5 * 7 + 4 * -5 + 3 * 1
    assert out[1, 3] == 4 * 1 + 3 * -2 + 7 * 3 + 7 * 2 + 4 * 1 + -10 * 0 + 4 * 7 + 3 * -5 + 2 * 1
    assert np.all(out[0] == 0) and np.all(out[2] == 0) and np.all(out[:, 0] == 0) and np.all(out[:, 4] == 0)


def test_gaussian_kernel():  # test_gaussian.py:9-33
    from lib.blur import gaussian

    k = gaussian.create_gaussian_kernel(5, 1.0)
    assert k.shape == (5, 5) and abs(1.0 - np.sum(k)) < 1e-7
    for i in range(3):
        for j in range(3):
            if i < 2:
                assert k[i, j] < k[i + 1, j]
            if j < 2:
                assert k[i, j] < k[i, j + 1]
            assert k[i, j] == k[4 - i, j] == k[i, 4 - j]
    with pytest.raises(ValueError):
        gaussian.create_gaussian_kernel(4, 1.0)


def test_detect_harris_corners():  # test_harris_detector.py:14-32
    from lib.harris import harris_detector as harris

    image = np.zeros((100, 200), dtype=float)
    image[25:76, 50:151] = 255  # cv.rectangle(background, (50, 25), (150, 75), 255, -1)
    expected = [(75, 50), (75, 150), (25, 50), (25, 150)]
    corners = harris.detect_harris_corners(image)
    assert len(corners) == 4
    for e in expected:  # the four corners carry exactly equal cornerness, so their order is a tie
        assert any(np.allclose(e, (c.y, c.x), atol=1.0) for c in corners)


def test_blur_then_detect_pipeline():
    """harris_detector's disabled visual test (:34-70) without the GUI: Gaussian blur -> ubyte -> corners."""
    from lib.blur import gaussian
    from lib.common import correlate
    from lib.harris import harris_detector as harris
    from oracle import front_end as fe
    from structure_from_motion_b200.scenes import make_image_pair

    img, *_ = make_image_pair(2, h=120, w=160)
    blurred = correlate.cross_correlate(img, gaussian.create_gaussian_kernel(3, 0.5)).astype(np.ubyte)
    want = fe.cross_correlate(img, gaussian.create_gaussian_kernel(3, 0.5)).astype(np.ubyte)
    assert np.abs(blurred.astype(int) - want.astype(int)).max() <= 1 and (blurred != want).mean() < 1e-3  # .5 boundaries
    corners = harris.detect_harris_corners(blurred, num_corners=30)
    xy, _, _, _ = fe.harris_corners_vectorised(blurred, 30)
    assert np.array_equal(np.array([[c.x, c.y] for c in corners]), xy)
