#!/bin/bash
# Round-2 profiling pass on the final build (one GPU):  gpurun --timeout 1500 -- 'bash tools/ncu_pass2.sh tag'
#   1. ncu launch list of the bench command (per-launch durations; shares must agree with bench.py's stage timers)
#   2. ncu --set full of the scoring kernel (k_score_auto) and of the tail kernels
tag=${1:-r2}; O=gpurun_out; mkdir -p $O
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fp32-variant --no-extras"
$BCMD > $O/plain_bench_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$tag.csv $BCMD > $O/ncu_launch_$tag.log 2>&1
echo "ncu launches rc=$?"
PCMD="python tools/run_two_view_once.py 2"
$PCMD > $O/plain_prof_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_score -c 1 -f -o $O/prof_score_$tag $PCMD > $O/ncu_prof_$tag.log 2>&1
echo "ncu full (k_score) rc=$?"
ncu --set full --clock-control none -k regex:'k_tail|k_finalise|k_fit_qr|k_screen' -c 7 -f -o $O/prof_tail_$tag $PCMD > $O/ncu_tail_$tag.log 2>&1
echo "ncu full (tail) rc=$?"
SFM_THR=1.5e-3 ncu --set full --clock-control none -k regex:k_score -c 1 -f -o $O/prof_score_thr1p5e-3_$tag python tools/run_score_once.py config3 auto 0 1 0 > $O/ncu_prof_hi_$tag.log 2>&1
echo "ncu full (k_score at thr 1.5e-3) rc=$?"
