#!/bin/bash
# one gpurun call: GPU parity suite, the default bench line, then ncu --set full of both passes on C4 and C3
set -x
mkdir -p gpurun_out
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest_$TAG.log
python bench.py > gpurun_out/r02_bench_$TAG.json 2> gpurun_out/r02_bench_$TAG.err; echo "bench rc=$?"
tail -2 gpurun_out/r02_bench_$TAG.err
for WL in c4:4096 c3:262144; do
  W=${WL%%:*}; N=${WL##*:}
  ncu --set full --clock-control none --import-source on -k regex:'lz_window|huff_lanes' --launch-skip 4 --launch-count 2 -f -o gpurun_out/prof_r2_${W}_$TAG \
    env SFB200_NO_OVERLAP=1 python tools/ab_bench.py --workloads $W --streams $N --variants "p:SFB200_NO_OVERLAP=1" --steps 1 --unique 256 > gpurun_out/prof_r2_${W}_$TAG.log 2>&1
  tail -2 gpurun_out/prof_r2_${W}_$TAG.log
done
ls -la gpurun_out/*.ncu-rep
