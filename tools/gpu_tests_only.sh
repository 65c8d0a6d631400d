#!/bin/bash
# usage: gpurun [--gpus N] --timeout 1200 -- 'bash tools/gpu_tests_only.sh tag'   the driver's GPU test command, log kept
tag=${1:-r2}; O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/gpus_$tag.txt
timeout 1000 python -m pytest tests/ -x -q -m gpu --timeout 300 > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$tag.log; tail -4 $O/pytest_$tag.log
