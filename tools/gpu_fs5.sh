#!/bin/bash
# usage: gpurun --timeout 900 -- 'bash tools/gpu_fs5.sh tag lib.so'   shipped: parity subset + sweep; headline A/B against a saved library
tag=$1; L=$2; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_selection.py tests/test_gpu_configs.py tests/test_gpu_batch.py -m gpu -q --timeout 200 > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$tag.log
bash tools/gpu_ab2.sh $L
timeout 300 python tools/thr_sweep.py 2>&1 | tee $O/thr_sweep_$tag.log
timeout 200 python tools/thr_sweep.py --shapes 2>&1 | grep -E "1.5e-06|1.5e-03" 
