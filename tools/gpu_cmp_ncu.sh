#!/bin/bash
# one gpurun call: compression tests + bench, then ncu --set full of the compression kernel
mkdir -p gpurun_out
TAG=${1:-cmpn}
python -m pytest tests -m gpu -x -q -k "compress" 2>&1 | tail -3
python tools/compress_bench.py --workload c2 --unique 2048 2>/dev/null | tee -a gpurun_out/r02_compress_$TAG.jsonl | cut -c1-420
python tools/compress_bench.py --workload c3 --unique 2048 2>/dev/null | tee -a gpurun_out/r02_compress_$TAG.jsonl | cut -c1-420
ncu --set full --clock-control none --import-source on -k regex:deflate_compress_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_r2_compress \
  python tools/compress_bench.py --workload c2 --streams 16384 --unique 1024 --steps 1 > gpurun_out/prof_r2_compress.log 2>&1
tail -2 gpurun_out/prof_r2_compress.log | cut -c1-300
ls -la gpurun_out/prof_r2_compress.ncu-rep
