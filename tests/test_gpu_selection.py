"""Selection parity in the regimes where summation order and dynamic range decide the winner
(lib/ransac/ransac.py:70-86, 96-108): noise-free data (every inlier score ~1e-30, the reference's own RANSAC test),
thresholds outside the fixed-point range (negative, +inf), and near-ties settled in the reference's list order (H1).
The models are the ORACLE's (engine.set_models), so the scores are bit-identical on both sides and the winner must be
the reference's, not merely an equivalent one."""
import random

import numpy as np
import pytest

from oracle import csed
from oracle import restatement as o
from structure_from_motion_b200 import two_view
from structure_from_motion_b200.scenes import make_scene

pytestmark = pytest.mark.gpu


def _oracle_run(K, x1, x2, thr, min_extra, agg, table):
    return o.ransac_essential(K, x1[:, 0], x1[:, 1], x2[:, 0], x2[:, 1], thr, min_extra, agg, len(table), table=table,
                              on_degenerate="skip", return_all=True, exact_sed=True)


@pytest.mark.parametrize("frac", [0.0, 0.25])
@pytest.mark.parametrize("agg", ["sum", "rms", "mean", "square"])
def test_noise_free_scene_selects_the_reference_winner(engine, agg, frac):
    """Noise-free correspondences, threshold 0.01 (test_epipolar.py:367-415).  Without outliers every hypothesis has an
    error of ~1e-28, 2^-80 of the threshold: K3 rescores them exactly (double-double), so the minimum-error hypothesis
    - not the earliest of a block of quantised zeros - wins, as in the reference.  With outliers the errors span
    thirty decades (the threshold is loose enough to admit gross outliers) and both accumulation paths are in play."""
    n, h, thr = 300, 96, 0.01
    K, x1, x2, *_ = make_scene(n, frac, seed=17, noise_px=0.0)
    rng = np.random.default_rng(5)
    table = np.stack([rng.choice(n, 8, replace=False) for _ in range(h)]).astype(np.int32)
    ref = _oracle_run(K, x1, x2, thr, 0, agg, table)
    engine.upload_pairs(x1, x2, K)
    engine.set_table(table)
    engine.set_models(ref["all_E"], ref["all_valid"])
    cnt, s1, s2, err = engine.score(thr, min_extra=0, aggregation=agg)
    if frac == 0.0:
        assert engine.rescored() == h  # every hypothesis went through the exact pass
    assert np.array_equal(cnt[ref["all_valid"]], ref["all_count"][ref["all_valid"]])
    fin = np.isfinite(ref["all_err"])
    assert np.array_equal(np.isfinite(err), fin)
    np.testing.assert_allclose(err[fin], ref["all_err"][fin], rtol=1e-12, atol=0)
    best = engine.get_best()
    assert best.index == ref["best_index"]
    assert abs(best.err - ref["error"]) <= 1e-12 * ref["error"]


def test_noise_free_winner_through_the_array_api(engine):
    """The same regime end to end (own fitter): the winner is the arg-min of the exactly summed errors of the GPU's
    own models, and its inlier list follows the reference order (samples, then the permutation tail)."""
    n, h, thr = 120, 64, 0.01
    K, x1, x2, *_ = make_scene(n, 0.0, seed=23, noise_px=0.0)
    random.seed(11)
    state = random.getstate()
    res = two_view.ransac_essential_arrays(K, x1, x2, thr, 0, "sum", h, engine=engine)
    E, valid = engine.get_models()
    random.setstate(state)
    perms = []
    perm = list(range(n))
    for _ in range(h):
        random.shuffle(perm)
        perms.append(list(perm))
    nxa, nya = o.k_normalise(x1[:, 0], x1[:, 1], K)
    nxb, nyb = o.k_normalise(x2[:, 0], x2[:, 1], K)
    errs = np.full(h, np.inf)
    for it in range(h):
        sed = csed.sed_exact_many(E[it], nxa, nya, nxb, nyb)
        order = np.asarray(perms[it])
        lst = np.concatenate([order[:8], order[8:][sed[order[8:]] <= thr]])
        errs[it] = sum(float(v) for v in sed[lst])
    assert res.best_index == int(np.argmin(errs))
    order = np.asarray(perms[res.best_index])
    assert res.inlier_indices[:8].tolist() == order[:8].tolist()
    assert res.inlier_indices.tolist() == [i for i in order if i in set(res.inlier_indices.tolist())]


@pytest.mark.parametrize("thr", [-1.0, float("inf"), float("nan"), 1e200])
def test_thresholds_outside_the_fixed_point_range(engine, thr):
    """ransac.py:70-75 accepts any float: a negative (or NaN) threshold gives no extra inliers, +inf makes every
    correspondence one.  (these used to be rejected.)"""
    n, h = 400, 40
    K, x1, x2, *_ = make_scene(n, 0.3, seed=4)
    rng = np.random.default_rng(1)
    table = np.stack([rng.choice(n, 8, replace=False) for _ in range(h)]).astype(np.int32)
    ref = _oracle_run(K, x1, x2, thr, 0, "rms", table)
    engine.upload_pairs(x1, x2, K)
    engine.set_table(table)
    engine.set_models(ref["all_E"], ref["all_valid"])
    cnt, s1, s2, err = engine.score(thr, min_extra=0, aggregation="rms")
    assert np.array_equal(cnt, ref["all_count"])
    assert (cnt == (0 if not thr > 0 else n - 8)).all()
    np.testing.assert_allclose(err, ref["all_err"], rtol=1e-12)
    assert engine.get_best().index == ref["best_index"]


def test_near_ties_follow_the_reference_summation_order(engine):
    """H1: models that differ in the last bits (the same eight correspondences re-drawn, LAPACK noise) have errors that
    agree to ~1e-15.  ransac.py:83 keeps the first one that is strictly smaller in the reference's own list-order
    arithmetic; the GPU's exactly rounded sums alone cannot know which that is, the host replay can."""
    n, thr, h = 600, 1.5e-6, 60
    K, x1, x2, *_ = make_scene(n, 0.3, seed=6)
    nxa, nya = o.k_normalise(x1[:, 0], x1[:, 1], K)
    nxb, nyb = o.k_normalise(x2[:, 0], x2[:, 1], K)
    ca, cb = np.stack([nxa, nya], 1), np.stack([nxb, nyb], 1)
    rng = np.random.default_rng(3)
    # a good model: the best of a few all-inlier-looking samples
    cands = [rng.choice(n, 8, replace=False) for _ in range(40)]
    fits = [o.eight_point(ca[c], cb[c]) for c in cands]
    good = int(np.argmax([(csed.sed_exact_many(e, nxa, nya, nxb, nyb) <= thr).sum() for e in fits]))
    base, e_base = cands[good], fits[good]
    table = np.empty((h, 8), dtype=np.int32)
    E = np.empty((h, 3, 3))
    for k in range(h):
        if k % 3 == 0:
            table[k] = rng.choice(n, 8, replace=False)
            E[k] = o.eight_point(ca[table[k]], cb[table[k]])
        else:  # the good model again, a few ulps away in one entry
            table[k] = base
            E[k] = e_base
            j = int(rng.integers(0, 8))
            v = E[k].reshape(9)[j]
            for _ in range(int(rng.integers(0, 4))):
                v = np.nextafter(v, np.inf if rng.integers(0, 2) else -np.inf)
            E[k].reshape(9)[j] = v
    hits = 0
    for agg in ("rms", "sum", "mean", "square"):
        # the reference's loop on these models: list order = samples, then the rest ascending; strict <
        want, want_err = -1, float("inf")
        for k in range(h):
            sed = csed.sed_exact_many(E[k], nxa, nya, nxb, nyb)
            keep = np.ones(n, dtype=bool)
            keep[table[k]] = False
            rest = np.nonzero(keep)[0]
            extra = rest[sed[rest] <= thr]
            if 5 <= len(extra):
                err = o.aggregate_error([float(v) for v in sed[np.concatenate([table[k], extra])]], agg)
                if err < want_err:
                    want, want_err = k, err
        engine.upload_pairs(x1, x2, K)
        engine.set_table(table)
        engine.set_models(E)
        engine.score(thr, min_extra=5, aggregation=agg, want_arrays=False)
        best = engine.get_best()
        ties, total = engine.near_ties(1e-12, 64)
        assert best.index in ties.tolist() and total == len(ties)

        def order_after(t):
            keep = np.ones(n, dtype=bool)
            keep[table[t]] = False
            return np.nonzero(keep)[0]

        t, e = two_view._resolve_near_ties(engine, ties, thr, agg, lambda t: table[t].astype(np.int64), order_after)
        assert t == want, (agg, t, want, best.index)
        assert e == want_err, agg  # the very same floating-point value: same scores, same summation order
        hits += int(total > 1)
        # and through the array API with this table (own fitter: only the tie-breaking machinery is exercised)
    assert hits == 4  # the scenario does produce near-ties
