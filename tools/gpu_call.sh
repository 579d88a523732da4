#!/bin/bash
# one gpurun call: GPU parity suite, a quick bench smoke, the default bench line
set -x
mkdir -p gpurun_out
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_$TAG.log
python bench.py --streams 4096 --unique 512 --shard-streams 16384 --steps 2 --warmup 3 --cpu-seconds 2 > gpurun_out/r02_bench_smoke_$TAG.json 2> gpurun_out/r02_bench_smoke_$TAG.err; echo "bench smoke rc=$?"
tail -3 gpurun_out/r02_bench_smoke_$TAG.err
cut -c1-600 gpurun_out/r02_bench_smoke_$TAG.json
( time python bench.py > gpurun_out/r02_bench_$TAG.json 2> gpurun_out/r02_bench_$TAG.err ) 2>&1 | tail -4; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_$TAG.err
python - <<'PY'
import json,sys,glob,os
p=sorted(glob.glob('gpurun_out/r02_bench_[!s]*.json'), key=os.path.getmtime)[-1]
d=json.loads(open(p).read().strip().splitlines()[-1])
print("C2", round(d["value"],1), "GB/s", round(d["ms_per_step"],2), "ms frac", round(d["roofline"]["frac"],4), d["roofline"]["passes_ms"])
print("e2e", d["e2e"]["value"], d["e2e"]["host_link"], "staging", d["e2e"]["device_staging_bytes"])
print("cpu", d.get("cpu_baseline",{}).get("value"))
print("sharded", d.get("sharded",{}).get("value"), d.get("sharded",{}).get("passes_ms"))
for k,v in d.get("shapes",{}).items(): print(k, round(v["value"],2), round(v["ms_per_step"],3), v.get("roofline",{}).get("passes_ms"))
PY
