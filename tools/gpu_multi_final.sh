#!/bin/bash
# usage: gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_multi_final.sh tag N [alltests]'   N-GPU validation: GPU tests + torchrun bench
tag=${1:-r2}; N=${2:-2}; O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/gpus_n${N}_$tag.txt; nproc >> $O/gpus_n${N}_$tag.txt
if [ "$3" == "alltests" ]; then T=tests; else T=tests/test_gpu_multi.py; fi
timeout 900 python -m pytest $T -m gpu -q --timeout 240 > $O/pytest_n${N}_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_n${N}_$tag.log; tail -4 $O/pytest_n${N}_$tag.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --no-cpu-baseline > $O/bench_n${N}_$tag.json 2> $O/bench_n${N}_$tag.err
echo "bench N=$N rc=$?"; tail -3 $O/bench_n${N}_$tag.err
python - <<PY
import json
d=json.load(open("$O/bench_n${N}_$tag.json"))
for k in ("value","ms_per_step","n_gpus","e2e","parity_multi","config3_strong","config5_strong","config4_pairs","clocks"):
    print(k, json.dumps(d.get(k))[:400])
PY
