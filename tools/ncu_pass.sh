#!/bin/bash
# The two ncu passes of /opt/skills/guides/B200_PROFILING.md on the current build (one GPU):
#   gpurun --timeout 1500 -- 'bash tools/ncu_pass.sh tag'
tag=${1:-r1}; O=gpurun_out; mkdir -p $O
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fp32-variant"
$BCMD > $O/plain_bench_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$tag.csv $BCMD > $O/ncu_launch_$tag.log 2>&1
echo "ncu launches rc=$?"
PCMD="python tools/run_score_once.py config3 screen 0 1 0"
$PCMD > $O/plain_prof_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_score -c 1 -f -o $O/prof_score_$tag $PCMD > $O/ncu_prof_$tag.log 2>&1
echo "ncu full rc=$?"
FCMD="python bench.py --workload frontend --steps 2 --warmup 3 --no-cpu-baseline"
$FCMD > $O/plain_frontend_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_frontend_$tag.csv $FCMD > $O/ncu_launch_frontend_$tag.log 2>&1
echo "ncu frontend launches rc=$?"
