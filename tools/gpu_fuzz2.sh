#!/bin/bash
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_fuzz2.sh tag'   long fuzz pass over every randomised checker
tag=${1:-r2}; O=gpurun_out; mkdir -p $O
for f in "fuzz_scorer 180" "fuzz_pipeline 150" "fuzz_batch 30" "fuzz_batch_two_view 30" "fuzz_list_api 40" "fuzz_front_end"; do
  set -- $f
  timeout 500 python tools/$1.py $2 > $O/$1_$tag.log 2>&1; echo "$1 $2 rc=$?"; grep -v " ok$" $O/$1_$tag.log | tail -4
done
