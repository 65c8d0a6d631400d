"""config 4 through distributed.PairPipeline.batch_two_view: contexts x chunk size (host buffers, whole path per pair)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_from_motion_b200 import _native  # noqa: E402
from structure_from_motion_b200.distributed import PairPipeline  # noqa: E402
from structure_from_motion_b200.scenes import make_scene  # noqa: E402

P, n, h = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 2000, 2000
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
for depth in ([int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else (1, 2, 3, 4)):
    pipe = PairPipeline(depth=depth)
    try:
        for chunk in ([int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else (128, 256, 512, 1024)):
            if chunk * depth > P:
                continue
            ts = []
            for r in range(4):
                t0 = time.perf_counter()
                out = pipe.batch_two_view(pa, pb, off, Ks, h, r, 1.5e-6, 10, "rms", chunk_pairs=chunk)
                ts.append(time.perf_counter() - t0)
            t = min(ts[1:])
            print(f"depth {depth} chunk {chunk:5d}: {t * 1e3:8.2f} ms  {P * n * h / t:.3e} evals/s  found {(out['best_index'] >= 0).sum()}")
    finally:
        pipe.close()
