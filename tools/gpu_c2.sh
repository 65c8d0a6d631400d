#!/bin/bash
# usage: gpurun --timeout 600 -- 'bash tools/gpu_c2.sh tag'   config-2 stage breakdown + launch list
tag=${1:-c2}; O=gpurun_out; mkdir -p $O
BCMD="python bench.py --workload config2 --steps 20 --warmup 3 --no-cpu-baseline --no-fp32-variant --no-extras"
timeout 200 $BCMD > $O/bench_c2_$tag.json 2> $O/bench_c2_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench_c2_$tag.json"))
print({k:d.get(k) for k in ("value","ms_per_step","stage_ms_per_step","gpu_launches")}, d["e2e"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c2_$tag.csv python bench.py --workload config2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fp32-variant --no-extras > $O/ncu_c2_$tag.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/launches_c2_$tag.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[-40:]: print(r[ki][:50], r[vi])
PY
