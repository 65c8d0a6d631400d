#!/bin/bash
# One gpurun call: GPU parity tests, smoke, both bench arms, then the two ncu passes of
# /opt/skills/guides/B200_PROFILING.md (launch list of the bench command; --set full of k_score).
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag] [noncu]'
tag=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu_$tag.txt 2>&1
python -m pytest tests -m gpu -x -q > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$tag.log
tail -3 $O/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_$tag.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$tag.json 2> $O/bench_ref_$tag.err; echo "ref rc=$?"
cat $O/bench_ref_$tag.json
python bench.py > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench rc=$?"
cat $O/bench_$tag.json; tail -5 $O/bench_$tag.err
python tools/time_score.py config3 > $O/time_score_$tag.log 2>&1; cat $O/time_score_$tag.log
if [ "$2" != "noncu" ]; then
  BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fp32-variant"
  $BCMD > $O/plain_bench_$tag.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$tag.csv $BCMD > $O/ncu_launch_$tag.log 2>&1
  echo "ncu launches rc=$?"
  PCMD="python tools/run_score_once.py config3 screen 0 1 0"
  $PCMD > $O/plain_prof_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_score -c 1 -f -o $O/prof_score_$tag $PCMD > $O/ncu_prof_$tag.log 2>&1
  echo "ncu full rc=$?"
fi
