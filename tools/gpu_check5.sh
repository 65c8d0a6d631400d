#!/bin/bash
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check5.sh tag'  (1 GPU: all GPU tests with a per-test timeout, smoke, bench, sweeps)
tag=${1:-r2}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 240 > $O/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$tag.log
tail -30 $O/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_$tag.log
timeout 600 python bench.py > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench_$tag.json"))
for k in ("value","ms_per_step","stage_ms_per_step","kernel_evals_per_s","gpu_launches","e2e","e2e_stream","cpu_baseline","config1_list_api","config2_latency","config3_strong","config5_strong","config4_pairs"):
    print(k, json.dumps(d.get(k))[:600])
print("roofline", json.dumps(d["roofline"])[:400])
PY
tail -5 $O/bench_$tag.err
timeout 200 python tools/time_fit.py > $O/time_fit_$tag.log 2>&1; cat $O/time_fit_$tag.log
timeout 300 python tools/thr_sweep.py > $O/thr_sweep_$tag.log 2>&1; cat $O/thr_sweep_$tag.log
