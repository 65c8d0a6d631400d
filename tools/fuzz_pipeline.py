"""One-off fuzz of the whole estimate (RANSAC E -> pose vote -> triangulation) against the numpy restatement of the
reference on the same sample table: random sizes, thresholds, outlier fractions, aggregation methods.
usage: python tools/fuzz_pipeline.py [cases]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import restatement as o  # noqa: E402
from structure_from_motion_b200 import _native, two_view  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(77)
eng = _native.get_engine(0)
bad = 0
for k in range(cases):
    n = int(rng.choice([40, 100, 333, 1000, 3000]))
    h = int(rng.choice([20, 64, 150, 400]))
    thr = float(10.0 ** rng.uniform(-7, -3.5))
    frac = float(rng.choice([0.0, 0.2, 0.5]))
    agg = ["rms", "sum", "mean", "square"][k % 4]
    min_extra = int(rng.choice([0, 5, 10]))
    noise = float(rng.choice([0.0, 0.1, 0.5, 1.0]))
    if noise == 0.0:  # noise-free: errors ~1e-28, the oracle must use the bit-faithful scalar scorer (slow): keep it small
        n, h = min(n, 100), min(h, 64)
    K, x1, x2, *_ = make_scene(n, frac, seed=500 + k, noise_px=noise)
    table = np.stack([rng.choice(n, 8, replace=False) for _ in range(h)]).astype(np.int32)
    msg = ""
    try:
        ref = o.ransac_essential(K, x1[:, 0], x1[:, 1], x2[:, 0], x2[:, 1], thr, min_extra, agg, h, table=table,
                                 on_degenerate="skip", exact_sed=(noise == 0.0))
    except ValueError:
        ref = None
    try:
        res = two_view.two_view_arrays(K, x1, x2, thr, min_extra, agg, h, sampler="table", table=table, on_degenerate="skip",
                                       engine=eng)
    except ValueError:
        res = None
    if (ref is None) != (res is None):
        msg = "one side found no model"
    elif ref is not None and noise == 0.0:
        # the models differ in their last bits (LAPACK vs QR), and at noise level that decides the winner: check the
        # GPU's choice against its OWN models - it must be the arg-min of the exactly summed errors
        from oracle import csed

        E, valid = eng.get_models()
        nxa, nya = o.k_normalise(x1[:, 0], x1[:, 1], K)
        nxb, nyb = o.k_normalise(x2[:, 0], x2[:, 1], K)
        best, best_err = -1, np.inf
        for it in range(h):
            if not valid[it]:
                continue
            sed = csed.sed_exact_many(E[it], nxa, nya, nxb, nyb)
            keep = np.ones(n, bool)
            keep[table[it]] = False
            extra = np.nonzero(keep & (sed <= thr))[0]
            if min_extra <= len(extra):
                err = o.aggregate_error([float(v) for v in sed[np.concatenate([table[it], extra])]], agg)
                if err < best_err:
                    best, best_err = it, err
        if res.ransac.best_index != best:
            msg = f"noise-free winner {res.ransac.best_index} vs {best}"
        elif abs(res.ransac.error - best_err) > 1e-12 * abs(best_err):
            msg = f"noise-free error {res.ransac.error} vs {best_err}"
    elif ref is not None:
        if res.ransac.best_index != ref["best_index"]:
            msg = f"winner {res.ransac.best_index} vs {ref['best_index']}"
        else:
            a, b = res.ransac.E.reshape(-1), ref["E"].reshape(-1)
            if np.abs(a / np.linalg.norm(a) - b / np.linalg.norm(b)).max() > 1e-6:
                msg = "E differs"
            nxa, nya = o.k_normalise(x1[:, 0], x1[:, 1], K)
            nxb, nyb = o.k_normalise(x2[:, 0], x2[:, 1], K)
            sed = o.sed_vectorised(nxa, nya, nxb, nyb, ref["E"])
            band = np.abs(sed - thr) <= 1e-9 * thr
            got = np.zeros(n, bool)
            got[res.ransac.inlier_indices] = True
            want = np.zeros(n, bool)
            want[ref["inlier_indices"]] = True
            if ((got != want) & ~band).any():
                msg = "inlier sets differ"
            # a winner supported by its 8 sample points only: its error is the fit's own rounding residual (1e-10 .. 1e-17
            # here), which two correct solvers reproduce to a few digits at best
            etol = 1e-9 if len(ref["inlier_indices"]) > 8 else 1e-6
            if abs(res.ransac.error - ref["error"]) > etol * abs(ref["error"]):
                msg = f"error {res.ransac.error} vs {ref['error']}"
            # apps/sfm.py:118-133 hands the RANSAC inlier LIST (samples first, ransac.py:76) to recover_r_t_from_e: its
            # position 0 - the correspondence np.count_nonzero never counts (eight_point.py:228-230) - is the first sample
            lst = ref["inlier_indices"]
            try:
                Rr, tr, mask_l, _ = o.recover_r_t_from_e(ref["E"], K, x1[lst, 0], x1[lst, 1], x2[lst, 0], x2[lst, 1])
                if not (np.allclose(res.R, Rr, atol=1e-6) and np.allclose(res.t, tr, atol=1e-6)):
                    msg = "pose differs"
                else:
                    pi = np.sort(lst[mask_l])
                    if not np.array_equal(pi, res.inlier_indices[res.passing]):
                        msg = "passing sets differ"
                    else:
                        X = o.triangulate_points(x1[pi, 0], x1[pi, 1], x2[pi, 0], x2[pi, 1], K, o.tmat(Rr, tr))
                        Xg = res.points[res.passing]
                        if len(X) and (np.linalg.norm(X - Xg, axis=1) / np.linalg.norm(X, axis=1)).max() > 1e-6:
                            msg = "points differ"
            except Exception as e:  # the reference raised in the pose stage: the GPU path must have raised too
                msg = f"oracle pose stage raised {type(e).__name__} but the GPU path returned"
    bad += bool(msg)
    print(f"{k:3d} n {n:5d} h {h:4d} thr {thr:8.2e} noise {noise:.1f} out {frac:.1f} {agg:6s} min_extra {min_extra:2d}  "
          f"{'no model' if ref is None else 'winner %4d inliers %4d' % (ref['best_index'], len(ref['inlier_indices']))}  {msg or 'ok'}")
print("mismatches:", bad)
sys.exit(1 if bad else 0)
