#!/bin/bash
# usage: gpurun --timeout 900 -- 'bash tools/gpu_fs2.sh tag [variant libs...]'   quick parity + sweep of auto/full at the two cliff thresholds
tag=${1:-fs}; shift; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_selection.py -m gpu -q --timeout 200 > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$tag.log
for r in 1 2; do
  echo "== shipped"; timeout 300 python tools/thr_sweep.py --cliff 2>&1 | grep -E "auto|full"
  for L in "$@"; do echo "== $L"; SFM_B200_LIB=$PWD/tools/bin/$L timeout 300 python tools/thr_sweep.py --cliff 2>&1 | grep -E "auto|full"; done
done
