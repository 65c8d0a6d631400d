#!/bin/bash
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_check4.sh tag'   (1 GPU: tests, smoke, K2 scaling-mode A/B, bench)
tag=${1:-r2}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$tag.log
tail -40 $O/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_$tag.log
for m in 0 1 2 3; do echo "scale mode $m"; SFM_B200_LIB=$PWD/tools/bin/libsfm_scale$m.so timeout 200 python tools/time_score.py config3 screen 2 16 2>&1 | tail -3; done | tee $O/scale_modes_$tag.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench rc=$?"
cut -c1-1500 $O/bench_$tag.json; echo; grep -o '"e2e".*' $O/bench_$tag.json | cut -c1-1800; tail -5 $O/bench_$tag.err
