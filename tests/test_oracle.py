"""The oracle pinned: oracle/restatement.py and oracle/sed_exact.c against the golden vectors
generated from the unmodified reference (tests/golden/make_golden.py), and — in the build
container only — live against /root/reference."""
import json
import os
import random

import numpy as np
import pytest

from oracle import csed
from oracle import restatement as o
from structure_from_motion_b200.scenes import make_scene

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def test_eight_point_fixture():
    d = load("eight_point_fixture.json")
    K = np.array(d["K"])
    p1, p2 = np.array(d["cam1_points"]), np.array(d["cam2_points"])
    f = o.eight_point(p1, p2)
    assert np.array_equal(f, np.array(d["F"]))  # same numpy calls => same bits
    na = np.stack(o.k_normalise(p1[:, 0], p1[:, 1], K), 1)
    nb = np.stack(o.k_normalise(p2[:, 0], p2[:, 1], K), 1)
    e = o.eight_point(na, nb)
    assert np.array_equal(e, np.array(d["E"]))
    np.testing.assert_almost_equal(np.array(d["E_opencv"]), e, decimal=5)  # test_epipolar.py:200-203
    np.testing.assert_almost_equal(np.array(d["F_opencv"]), f, decimal=5)  # test_epipolar.py:188-191
    R1, R2, t1 = o.recover_all_r_t(e.copy())
    assert np.array_equal(R1, np.array(d["R1"])) and np.array_equal(R2, np.array(d["R2"]))
    assert np.array_equal(t1, np.array(d["t1"]))
    R, t, mask, _ = o.recover_r_t(na[:, 0], na[:, 1], nb[:, 0], nb[:, 1], e.copy())
    assert np.array_equal(R, np.array(d["R"])) and np.array_equal(t, np.array(d["t"]))
    assert mask.tolist() == d["mask"] == list(range(8))  # test_epipolar.py:240-244
    tn = t / np.linalg.norm(t)
    np.testing.assert_allclose(np.array(d["expected_t_direction"]), tn, atol=1e-5, rtol=0)  # :256-262
    np.testing.assert_allclose(np.array(d["expected_R"]), R, atol=1e-6)
    for k in range(8):  # test_epipolar.py:501-515
        s = o.sed_scalar(na[k, 0], na[k, 1], nb[k, 0], nb[k, 1], np.array(d["E_opencv"]))
        assert s == d["sed_under_E_opencv"][k] and s < 1e-20


def test_ransac_known_answer():
    d = load("ransac_known_answer.json")
    pa, pb = np.array(d["pts_a"]), np.array(d["pts_b"])
    random.seed(d["seed"])
    r = o.ransac_essential(np.array(d["K"]), pa[:, 0], pa[:, 1], pb[:, 0], pb[:, 1], d["threshold"],
                           None, d["method"], None, exact_sed=True)
    assert np.array_equal(r["E"], np.array(d["E"]))
    assert r["inlier_indices"].tolist() == d["inlier_indices"]
    assert list(random.getstate()[1]) == d["rng_state_after"]


def test_config1_known_answer():
    d = load("config1_known_answer.json")
    K, x1, x2, *_ = make_scene(**d["scene"])
    random.seed(d["seed"])
    r = o.ransac_essential(K, x1[:, 0], x1[:, 1], x2[:, 0], x2[:, 1], d["threshold"], d["min_extra"],
                           d["method"], d["max_iterations"])
    assert r["best_index"] == d["best_index"] == 87
    assert np.array_equal(r["E"], np.array(d["E"]))
    assert r["inlier_indices"].tolist() == d["inlier_indices"] and len(d["inlier_indices"]) == 23
    assert abs(r["error"] - d["error"]) <= 1e-12 * d["error"]


def test_sampler_known_answer():
    d = load("sampler_known_answer.json")
    random.seed(d["seed"])
    t = o.python_shuffle_table(d["n"], 3)
    assert t.tolist() == d["rows"]
    assert t[0].tolist() == [910, 516, 275, 950, 612, 970, 52, 75]


def test_degenerate_fixture():
    d = load("degenerate_fixture.json")
    assert d["raises"]
    with pytest.raises(o.OracleEightPointError):
        o.eight_point(np.array(d["cam1_points"]), np.array(d["cam2_points"]))


def test_triangulation_known_answer():
    d = load("triangulation_known_answer.json")
    X = o.triangulate_one(*d["feature_a"], *d["feature_b"], np.array(d["P1"]), np.array(d["P2"]))
    assert np.array_equal(X, np.array(d["reference"]))
    np.testing.assert_allclose(np.array(d["expected"]), X, atol=1e-10, rtol=0)  # test_epipolar.py:494-496


def test_sed_vectors_bit_exact():
    rows = load("sed_vectors.json")
    for r in rows:
        E = np.array(r["E"])
        assert o.sed_scalar(r["xa"], r["ya"], r["xb"], r["yb"], E) == r["sed"]
        c = csed.sed_exact_many(E, [r["xa"]], [r["ya"]], [r["xb"]], [r["yb"]])[0]
        assert c == r["sed"], (c, r["sed"])  # the C scorer reproduces numpy's evaluation order


def test_pose_known_answer():
    d = load("pose_known_answer.json")
    K, pa, pb = np.array(d["K"]), np.array(d["pts_a"]), np.array(d["pts_b"])
    R, t, mask, _ = o.recover_r_t_from_e(np.array(d["E"]), K, pa[:, 0], pa[:, 1], pb[:, 0], pb[:, 1])
    assert np.array_equal(R, np.array(d["R"])) and np.array_equal(t, np.array(d["t"]))
    assert mask.tolist() == d["mask"]
    X = o.triangulate_points(pa[:, 0], pa[:, 1], pb[:, 0], pb[:, 1], K, o.tmat(R, t))
    assert np.array_equal(X, np.array(d["X"]))


def test_c_scorer_batch_matches_restatement():
    K, x1, x2, *_ = make_scene(700, 0.4, seed=3)
    nxa, nya = o.k_normalise(x1[:, 0], x1[:, 1], K)
    nxb, nyb = o.k_normalise(x2[:, 0], x2[:, 1], K)
    rng = np.random.default_rng(0)
    table = np.stack([rng.choice(700, 8, replace=False) for _ in range(40)]).astype(np.int32)
    E = np.stack([o.eight_point(np.stack([nxa[s], nya[s]], 1), np.stack([nxb[s], nyb[s]], 1)) for s in table])
    cnt, s1, s2 = csed.score_batch(E, nxa, nya, nxb, nyb, 1.5e-6, table=table, nthreads=3)
    for h in range(40):
        sed = np.array([o.sed_scalar(nxa[i], nya[i], nxb[i], nyb[i], E[h]) for i in range(700)])
        samp = np.zeros(700, bool)
        samp[table[h]] = True
        extra = (sed <= 1.5e-6) & ~samp
        assert cnt[h] == extra.sum()
        np.testing.assert_allclose(s1[h], sed[samp | extra].sum(), rtol=1e-13)
        np.testing.assert_allclose(s2[h], (sed[samp | extra] ** 2).sum(), rtol=1e-13)


@pytest.mark.reference
def test_restatement_matches_live_reference():
    """Build container only: the unmodified reference, run here, against the restatement."""
    from oracle import reference_shims

    ref = reference_shims.load()
    F, M = ref.feature.Feature, ref.matching.Match
    K, x1, x2, *_ = make_scene(150, 0.3, seed=2)
    fa = [F(x=float(p[0]), y=float(p[1])) for p in x1]
    fb = [F(x=float(p[0]), y=float(p[1])) for p in x2]
    ms = [M(a_index=i, b_index=i) for i in range(150)]
    for method in ("sum", "square", "mean", "rms"):
        random.seed(9)
        e, pairs = ref.epipolar_ransac.estimate_essential_mat_with_ransac(
            K, fa, fb, ms, 1.5e-6, min_num_extra_inliers=4,
            error_aggregation_method=ref.ransac.ErrorAggregationMethod(method), max_iterations=40)
        st = random.getstate()
        random.seed(9)
        r = o.ransac_essential(K, x1[:, 0], x1[:, 1], x2[:, 0], x2[:, 1], 1.5e-6, 4, method, 40, exact_sed=True)
        assert random.getstate() == st
        assert np.array_equal(e, r["E"])
        assert [(p[0].x, p[1].x) for p in pairs] == [(x1[i, 0], x2[i, 0]) for i in r["inlier_indices"]]
    inl = r["inlier_indices"]
    R, t, mask = ref.eight_point.recover_r_t_from_e(e.copy(), K, [fa[i] for i in inl], [fb[i] for i in inl])
    R2, t2, mask2, _ = o.recover_r_t_from_e(e.copy(), K, x1[inl, 0], x1[inl, 1], x2[inl, 0], x2[inl, 1])
    assert np.array_equal(R, R2) and np.array_equal(t, t2) and np.array_equal(mask, mask2)
    X = ref.triangulation.triangulate_points([fa[i] for i in inl], [fb[i] for i in inl], K,
                                             ref.transforms.Transform3D.from_rmat_t(R, t))
    X2 = o.triangulate_points(x1[inl, 0], x1[inl, 1], x2[inl, 0], x2[inl, 1], K, o.tmat(R, t))
    assert np.array_equal(X, X2)
