#!/bin/bash
# one gpurun call: default bench line (all shapes), then the ncu launch list of the same command
# (headline only: the other shapes would take minutes under the profiler)
mkdir -p gpurun_out
TAG=${1:-final}
( time python bench.py > gpurun_out/r02_bench_$TAG.json 2> gpurun_out/r02_bench_$TAG.err ) 2>&1 | tail -3; echo "bench rc=$?"
( time python bench.py --impl reference > gpurun_out/r02_bench_ref_$TAG.json 2> gpurun_out/r02_bench_ref_$TAG.err ) 2>&1 | tail -3
cut -c1-300 gpurun_out/r02_bench_ref_$TAG.json
python bench.py --steps 2 --warmup 3 --no-shapes --no-cpu --no-e2e > gpurun_out/r02_bench_plain_$TAG.json 2> /dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c2_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-shapes --no-cpu --no-e2e > gpurun_out/r02_ncu_$TAG.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r02_launches_c2_$TAG.csv
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_$TAG.json').read().strip().splitlines()[-1])
print("C2", round(d["value"],1), "GB/s", round(d["ms_per_step"],2), "ms frac", round(d["roofline"]["frac"],4), d["roofline"]["passes_ms"], d["clocks"])
print("e2e", d["e2e"]["value"], d["e2e"]["host_link"])
print("cpu", d.get("cpu_baseline",{}).get("value"))
print("sharded", d.get("sharded",{}).get("value"), d.get("sharded",{}).get("passes_ms"), d.get("sharded",{}).get("clocks"))
for k,v in d.get("shapes",{}).items(): print(k, round(v["value"],2), round(v["ms_per_step"],3), v.get("steps"), v.get("clocks"))
PY
