#!/bin/bash
# usage: gpurun --timeout 600 -- 'bash tools/gpu_fit.sh tag libs...'   K1 timing A/B + fitter parity
tag=${1:-fit}; shift; O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q --timeout 200 -k "fit or golden or degenerate or essential" > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$tag.log
for r in 1 2; do
  echo "== shipped"; timeout 100 python tools/time_fit.py 2>&1 | head -2
  for L in "$@"; do echo "== $L"; SFM_B200_LIB=$PWD/tools/bin/$L timeout 100 python tools/time_fit.py 2>&1 | head -2; done
done
