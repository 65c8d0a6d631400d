#!/bin/bash
# usage: gpurun --timeout 600 -- 'bash tools/gpu_ab2.sh libs...'   K2 headline timing, shipped vs variants, alternating, 3 rounds
for r in 1 2 3; do
  echo "== shipped"; timeout 100 python tools/time_score.py config3 auto 2 16 2>/dev/null | head -2
  for L in "$@"; do echo "== $L"; SFM_B200_LIB=$PWD/tools/bin/$L timeout 100 python tools/time_score.py config3 auto 2 16 2>/dev/null | head -2; done
done
