"""One-off fuzz of the stages in front of the hot path against oracle/front_end.py: random images (uint8 / float64),
feature positions (inside, on the border, outside, half-pixel), window sizes 1..15, NCC / SSD, every validation
combination; Harris with random block sizes and k on random uint8 images.  usage: python tools/fuzz_front_end.py [cases]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import front_end as fe  # noqa: E402
from structure_from_motion_b200 import _native  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(4242)
eng = _native.get_engine(0)
bad = 0
for k in range(cases):
    h, w = int(rng.integers(8, 70)), int(rng.integers(8, 90))
    u8 = bool(k % 3)
    img_a = rng.integers(0, 256, (h, w)).astype(np.uint8)
    img_b = np.clip(img_a.astype(int) + rng.integers(-20, 21, (h, w)), 0, 255).astype(np.uint8)
    if not u8:
        img_a, img_b = img_a / 255.0 + rng.random((h, w)) * 1e-3, img_b / 255.0
    if k % 7 == 0:
        img_a[:] = img_a.flat[0]  # no texture: NCC denominator 0
    na, nb = int(rng.integers(1, 40)), int(rng.integers(1, 40))
    fa = np.stack([rng.uniform(-3, w + 3, na), rng.uniform(-3, h + 3, na)], 1)
    fb = np.stack([rng.uniform(-3, w + 3, nb), rng.uniform(-3, h + 3, nb)], 1)
    fa[::2] = np.floor(fa[::2])
    fb[::3] = np.floor(fb[::3]) + 0.5
    kind = "ncc" if k % 2 else "ssd"
    win = int(rng.integers(1, 16))
    ratio, cross = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    thr = float(rng.uniform(0.3, 1.1))
    msg = ""
    with np.errstate(all="ignore"):
        S_o = fe.score_matrix(img_a, img_b, fa, fb, kind, win)
    bb, bs, keep, S = eng.match_brute_force(img_a, img_b, fa, fb, kind=kind, window=win, ratio_test=ratio, crosscheck=cross,
                                            ratio_threshold=thr, want_scores=True)
    if not np.array_equal(np.isfinite(S), np.isfinite(S_o)) or not np.array_equal(S == 2.0, S_o == 2.0):
        msg = "outside / no-texture pattern differs"
    else:
        fin = np.isfinite(S_o)
        tol = 1e-12 if kind == "ncc" else (0.0 if u8 else 1e-14 * max(1.0, np.abs(S_o[fin]).max() if fin.any() else 1.0))
        if fin.any() and np.abs(S[fin] - S_o[fin]).max() > tol:
            msg = f"scores differ by {np.abs(S[fin] - S_o[fin]).max():.3e}"
    if not msg:
        with np.errstate(all="ignore"):
            want = fe.match_from_scores(S, ratio, cross, thr)  # selection must be exact given the GPU's own matrix
        got = [(int(a), int(bb[a]), float(bs[a])) for a in np.flatnonzero(keep)]
        if got != want:
            msg = "selection differs"
    # Harris on the uint8 version of image b
    imgh = img_b if u8 else np.clip(np.round(img_b * 255), 0, 255).astype(np.uint8)
    bs_, kk, num = int(rng.integers(1, 6)), float(rng.choice([0.04, 0.06, 0.15])), int(rng.integers(1, 60))
    if not msg and h - bs_ > 0 and w - bs_ > 0:
        xy, sc, extra = eng.harris_corners(imgh, num, bs_, kk, want_cornerness=True)
        xy_o, sc_o, cim_o, _ = fe.harris_corners_vectorised(imgh, num, bs_, kk)
        raw = fe.cornerness_image_vectorised(imgh, bs_, kk)
        raw[raw < 0] = 0.0
        seq = raw.copy()
        fe.non_max_suppress(seq)
        if not (np.array_equal(extra["cornerness"], cim_o) and np.array_equal(cim_o, seq)):
            msg = "cornerness / suppression differs"
        elif not (np.array_equal(xy, xy_o) and np.array_equal(sc, sc_o)):
            msg = "corner list differs"
    bad += bool(msg)
    print(f"{k:3d} {h:2d}x{w:2d} {'u8 ' if u8 else 'f64'} {kind} w {win:2d} na {na:2d} nb {nb:2d} ratio {int(ratio)} cross {int(cross)} "
          f"harris bs {bs_} num {num:2d}  {msg or 'ok'}")
print("mismatches:", bad)
sys.exit(1 if bad else 0)
