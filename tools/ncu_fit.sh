#!/bin/bash
# usage: gpurun --timeout 600 -- 'bash tools/ncu_fit.sh tag'   ncu --set full of K1 with source
tag=${1:-fit}; O=gpurun_out; mkdir -p $O
PCMD="python tools/run_two_view_once.py 2"
timeout 120 $PCMD > $O/plain_fit_$tag.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fit_qr -c 2 -f -o $O/prof_fit_$tag $PCMD > $O/ncu_fit_$tag.log 2>&1
echo "ncu rc=$?"
