#!/bin/bash
# usage: gpurun --timeout 900 -- 'bash tools/gpu_ea.sh lib.so'   variant lib: parity + headline A/B + sweep
L=$1; O=gpurun_out; mkdir -p $O
SFM_B200_LIB=$PWD/tools/bin/$L timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_selection.py -m gpu -q --timeout 200 > $O/pytest_$L.log 2>&1; echo "pytest($L) rc=$?"; tail -3 $O/pytest_$L.log
bash tools/gpu_ab2.sh $L
echo "== sweep $L"; SFM_B200_LIB=$PWD/tools/bin/$L timeout 300 python tools/thr_sweep.py 2>&1
