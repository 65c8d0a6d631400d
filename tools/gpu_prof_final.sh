#!/bin/bash
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_prof_final.sh tag'   final profiling pass: ncu_pass2 + threshold sweep + fuzzers
tag=${1:-r2}; O=gpurun_out; mkdir -p $O
bash tools/ncu_pass2.sh $tag
timeout 300 python tools/thr_sweep.py > $O/thr_sweep_$tag.log 2>&1; cat $O/thr_sweep_$tag.log
timeout 200 python tools/fuzz_pipeline.py 40 > $O/fuzz_pipeline_$tag.log 2>&1; echo "fuzz_pipeline rc=$?"; tail -1 $O/fuzz_pipeline_$tag.log
timeout 200 python tools/fuzz_batch_two_view.py 10 > $O/fuzz_b2v_$tag.log 2>&1; echo "fuzz_batch_two_view rc=$?"; tail -1 $O/fuzz_b2v_$tag.log
timeout 200 python tools/fuzz_scorer.py 30 > $O/fuzz_scorer_$tag.log 2>&1; echo "fuzz_scorer rc=$?"; tail -1 $O/fuzz_scorer_$tag.log
