#!/bin/bash
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_check7.sh tag'   tests + thr sweep + fit register variants + pipeline fuzz
tag=${1:-r2}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 240 > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$tag.log; tail -4 $O/pytest_$tag.log
timeout 300 python tools/thr_sweep.py > $O/thr_sweep_$tag.log 2>&1; cat $O/thr_sweep_$tag.log
timeout 300 python tools/fuzz_pipeline.py 60 > $O/fuzz_pipeline_$tag.log 2>&1; echo "fuzz_pipeline rc=$?"; grep -v " ok$" $O/fuzz_pipeline_$tag.log | tail -5
timeout 100 python tools/time_fit.py 2>&1 | head -2
for R in 136 144 152; do echo "maxnreg $R"; SFM_B200_LIB=$PWD/tools/bin/libsfm_fitreg$R.so timeout 100 python tools/time_fit.py 2>&1 | head -2; done
