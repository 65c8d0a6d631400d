"""Randomised parity sweeps (tools/fuzz_scorer.py, tools/fuzz_pipeline.py) in a size that fits the test budget: random
problem sizes (8 .. 20 000 correspondences, 1 .. 4 097 hypotheses), thresholds over ten decades, outlier fractions
0 .. 0.9, every scoring variant and aggregation method — counts bit-equal to the exact C scorer, sums within 1e-12,
same winner; the whole estimate (winner, E, inliers, pose, points, no-model cases) against the numpy restatement;
the matcher (scores, selection, validations) and the Harris detector against oracle/front_end.py; ragged pair batches
(empty pairs, fewer than eight correspondences, every selection mode) against the single-pair pipeline; the drop-in list API (E, inlier pairs in order, global RNG state, exception types)
against the restatement."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,cases", [("fuzz_scorer.py", 25), ("fuzz_pipeline.py", 12), ("fuzz_front_end.py", 30), ("fuzz_batch.py", 15), ("fuzz_list_api.py", 24)])
def test_fuzz(script, cases):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script), str(cases)], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "mismatches: 0" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
