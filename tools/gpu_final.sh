#!/bin/bash
# The driver's round-end sequence on one GPU: GPU tests, smoke, reference arm, native arm.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_final.sh tag'
tag=${1:-r2}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu_$tag.txt 2>&1; nproc >> $O/gpu_$tag.txt
if [ "$2" != "notests" ]; then timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$tag.log; tail -4 $O/pytest_$tag.log; fi
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$tag.log
t0=$(date +%s); python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_ref_$tag.json 2> $O/bench_ref_$tag.err; echo "ref rc=$? wall $(( $(date +%s) - t0 )) s"; cut -c1-400 $O/bench_ref_$tag.json
t0=$(date +%s); python bench.py --gpus 1 --steps 20 --warmup 3 > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"
python - <<PY
import json
d=json.load(open("$O/bench_$tag.json"))
for k in ("value","ms_per_step","stage_ms_per_step","kernel_evals_per_s","gpu_launches","e2e","e2e_stream","fp32_prefilter","clocks","config1_list_api","config2_latency","config3_strong","config5_strong","config4_pairs"):
    print(k, json.dumps(d.get(k))[:330])
print("roofline", json.dumps(d["roofline"])[:300]); print("cpu_baseline", json.dumps(d["cpu_baseline"])[:600])
PY
