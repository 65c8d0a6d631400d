#!/bin/bash
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_check3.sh tag'   (1 GPU: tests, smoke, bench, fit variants)
tag=${1:-r2}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$tag.log
tail -25 $O/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_$tag.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench rc=$?"
cut -c1-2600 $O/bench_$tag.json; echo; grep -o '"config3_strong.*' $O/bench_$tag.json | cut -c1-1800; tail -5 $O/bench_$tag.err
timeout 200 python tools/time_fit.py > $O/time_fit_$tag.log 2>&1; cat $O/time_fit_$tag.log
for v in 7; do SFM_B200_LIB=$PWD/tools/bin/libsfm_fit$v.so timeout 200 python tools/time_fit.py > $O/time_fit${v}_$tag.log 2>&1; echo "MINB $v"; cat $O/time_fit${v}_$tag.log; done
