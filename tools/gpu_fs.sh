#!/bin/bash
# usage: gpurun --timeout 900 -- 'bash tools/gpu_fs.sh tag'   two-sided body with shared-memory gathers: parity + A/B
tag=${1:-fs}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_selection.py tests/test_gpu_configs.py -m gpu -q --timeout 200 > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_$tag.log
timeout 300 python tools/thr_sweep.py > $O/thr_sweep_$tag.log 2>&1; cat $O/thr_sweep_$tag.log
echo "== base sweep"; SFM_B200_LIB=$PWD/tools/bin/base.so timeout 300 python tools/thr_sweep.py 2>&1 | grep -E "auto|full"
for r in 1 2 3; do
  echo "== new";  timeout 100 python tools/time_score.py config3 auto 2 16 | head -1
  echo "== base"; SFM_B200_LIB=$PWD/tools/bin/base.so timeout 100 python tools/time_score.py config3 auto 2 16 | head -1
done
timeout 200 python tools/fuzz_pipeline.py 30 > $O/fuzz_pipeline_$tag.log 2>&1; echo "fuzz_pipeline rc=$?"; grep -v " ok$" $O/fuzz_pipeline_$tag.log | tail -5
