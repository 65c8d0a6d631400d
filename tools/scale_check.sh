#!/bin/bash
# Multi-GPU bench lines on one box: usage  gpurun --gpus N --timeout 900 -- 'bash tools/scale_check.sh N tag'
N=${1:-2}; tag=${2:-r1}; O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/gpus_${N}_$tag.txt
run() { # name, extra args
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --no-cpu-baseline "$@" > $O/bench_n${N}_${name}_$tag.json 2> $O/bench_n${N}_${name}_$tag.err
  echo "$name rc=$?"; tail -1 $O/bench_n${N}_${name}_$tag.json | cut -c1-600
}
run config3 --steps 20 --warmup 3
run config5 --workload config5 --steps 3 --warmup 3 --no-e2e --no-fp32-variant
run config4 --workload config4 --steps 10 --warmup 3
