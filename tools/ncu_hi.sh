#!/bin/bash
# usage: gpurun --timeout 600 -- 'bash tools/ncu_hi.sh tag'   ncu --set full of the scoring kernel at a 17 % inlier rate
tag=${1:-hi}; O=gpurun_out; mkdir -p $O
SFM_THR=1.5e-3 timeout 120 python tools/run_score_once.py config3 auto 0 1 0 > $O/plain_hi_$tag.log 2>&1 &&
SFM_THR=1.5e-3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_score -c 1 -f -o $O/prof_score_thr1p5e-3_$tag python tools/run_score_once.py config3 auto 0 1 0 > $O/ncu_prof_hi_$tag.log 2>&1
echo "ncu full (k_score at thr 1.5e-3) rc=$?"
