#!/bin/bash
# one gpurun call: compression tests + bench (dynamic and fixed-only) on C2 and C3 plaintexts
mkdir -p gpurun_out
TAG=${1:-cmp}
python -m pytest tests -m gpu -x -q -k "compress" 2>&1 | tail -3
for WL in c2 c3; do
  python tools/compress_bench.py --workload $WL --unique 2048 2>/dev/null | tee -a gpurun_out/r02_compress_$TAG.jsonl | cut -c1-420
  SFB200_COMPRESS_FIXED=1 python tools/compress_bench.py --workload $WL --unique 2048 2>/dev/null | tee -a gpurun_out/r02_compress_$TAG.jsonl | cut -c1-420
done
