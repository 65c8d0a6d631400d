#!/bin/bash
# A/B of a variant library against the shipped one on the K2 timing (alternating, 3 rounds)
O=gpurun_out; mkdir -p $O
for r in 1 2 3; do
  echo "== shipped"; timeout 100 python tools/time_score.py config3 auto 2 16 | head -1
  echo "== variant $1"; SFM_B200_LIB=$PWD/tools/bin/$1 timeout 100 python tools/time_score.py config3 auto 2 16 | head -1
done
