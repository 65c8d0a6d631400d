#!/bin/bash
# usage: gpurun [--gpus 2] --timeout 1500 -- 'bash tools/gpu_check6.sh tag [N]'
tag=${1:-r2}; N=${2:-1}; O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/gpus_$tag.txt; nproc >> $O/gpus_$tag.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 240 > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$tag.log
tail -6 $O/pytest_$tag.log
timeout 400 python bench.py --no-cpu-baseline --no-fp32-variant > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench_$tag.json"))
for k in ("value","ms_per_step","stage_ms_per_step","kernel_evals_per_s","gpu_launches","e2e","e2e_stream","config2_latency","config4_pairs"):
    print(k, json.dumps(d.get(k))[:400])
PY
if [ "$N" != "1" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --no-cpu-baseline > $O/bench_n${N}_$tag.json 2> $O/bench_n${N}_$tag.err
  echo "bench N=$N rc=$?"; tail -3 $O/bench_n${N}_$tag.err
  python - <<PY
import json
d=json.load(open("$O/bench_n${N}_$tag.json"))
for k in ("value","ms_per_step","n_gpus","e2e","parity_multi","config3_strong","config5_strong","config4_pairs"):
    print(k, json.dumps(d.get(k))[:400])
PY
fi
timeout 300 python tools/pipe_depth.py 2048 > $O/pipe_depth_$tag.log 2>&1; cat $O/pipe_depth_$tag.log
