#!/bin/bash
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_fuzz.sh tag'
tag=${1:-r2}; O=gpurun_out; mkdir -p $O
for f in fuzz_scorer fuzz_pipeline fuzz_batch fuzz_batch_two_view fuzz_list_api fuzz_front_end; do
  timeout 400 python tools/$f.py > $O/${f}_$tag.log 2>&1; echo "$f rc=$?"; grep -v " ok$" $O/${f}_$tag.log | tail -6
done
SFM_THR=1.5e-3 ncu --set full --clock-control none -k regex:k_score -c 1 -f -o $O/prof_score_hi2_$tag python tools/run_score_once.py config3 auto 0 1 0 > $O/ncu_prof_hi2_$tag.log 2>&1; echo "ncu hi rc=$?"
