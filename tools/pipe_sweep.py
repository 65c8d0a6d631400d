"""config 4 (512 pairs x 2000 x 2000, host buffers): PairPipeline depth / chunking sweep.  usage: python tools/pipe_sweep.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.distributed import PairPipeline  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

P, n, h = 512, 2000, 2000
base = make_scene(n, 0.4, seed=0)
rng = np.random.default_rng(0)
pa = _native.pinned_empty((P * n, 2))
pb = _native.pinned_empty((P * n, 2))
for p in range(P):
    perm = rng.permutation(n)
    pa[p * n:(p + 1) * n] = base[1][perm]
    pb[p * n:(p + 1) * n] = base[2][perm]
off = np.arange(P + 1, dtype=np.int64) * n
Ks = np.stack([base[0]] * P)
eng = _native.get_engine(0)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for s in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


print("single call            %.3f ms" % timeit(lambda: eng.batch_ransac(pa, pb, off, Ks, h, 1, 1.5e-6, 10, "rms")))
for depth in (2, 3):
    pipe = PairPipeline(depth=depth)
    for chunk in (None, 256, 128, 64, 32):
        ms = timeit(lambda: pipe.batch_ransac(pa, pb, off, Ks, h, 1, 1.5e-6, 10, "rms", chunk_pairs=chunk))
        print("depth %d chunk %-6s   %.3f ms" % (depth, chunk, ms))
    pipe.close()
