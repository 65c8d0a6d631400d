#!/bin/bash
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_prof_tail.sh tag'
tag=${1:-r2}; O=gpurun_out; mkdir -p $O
python tools/run_two_view_once.py 3 > $O/plain_two_view_$tag.log 2>&1 && cat $O/plain_two_view_$tag.log
ncu --set full --clock-control none --import-source on -k regex:'k_tail|k_finalise|k_fit_qr|k_sample|k_screen' -c 14 -f -o $O/prof_tail_$tag python tools/run_two_view_once.py 2 > $O/ncu_tail_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 $O/ncu_tail_$tag.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_two_view_$tag.csv python tools/run_two_view_once.py 3 > /dev/null 2>&1; echo "launch list rc=$?"
