#!/bin/bash
# Round-2 GPU check: parity tests (multi-rank included when the lease has >= 2 GPUs), smoke, bench arms, K2 variants.
# usage: gpurun [--gpus N] --timeout 2400 -- 'bash tools/gpu_check2.sh tag [N]'
tag=${1:-r2}; N=${2:-1}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu_$tag.txt 2>&1
nproc >> $O/gpu_$tag.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$tag.log
tail -15 $O/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_$tag.log
timeout 600 python bench.py > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench rc=$?"
cut -c1-3000 $O/bench_$tag.json; tail -5 $O/bench_$tag.err
if [ "$N" != "1" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --no-cpu-baseline > $O/bench_n${N}_$tag.json 2> $O/bench_n${N}_$tag.err
  echo "bench N=$N rc=$?"; cut -c1-3000 $O/bench_n${N}_$tag.json; tail -5 $O/bench_n${N}_$tag.err
fi
timeout 300 python tools/time_score.py config3 screen 2 16 > $O/time_score_$tag.log 2>&1
timeout 300 python tools/time_score.py config3 screen 4 8 >> $O/time_score_$tag.log 2>&1; cat $O/time_score_$tag.log
