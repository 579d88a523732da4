#!/bin/bash
# round 2, GPU call 1: parity of the segment resolve kernel, v1 vs v2 timing on C2/C3/C4, ncu captures
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest1.log
rm -f gpurun_out/r02_ab1.jsonl
python tools/ab_bench.py --workloads c2,c3,c4 --variants "v2:;v1:SFB200_LZ_V1=1" --steps 5 --unique 4096 --out gpurun_out/r02_ab1.jsonl 2> gpurun_out/r02_ab1.err | cut -c1-400
NCU="ncu --set full --import-source on --clock-control none"
timeout 300 $NCU -k regex:"huff_lanes_kernel|lz_resolve_kernel" -c 2 -f -o gpurun_out/r02_c2_seg python tools/ab_bench.py --workloads c2 --streams 32768 --steps 0 --variants "v2:" > gpurun_out/r02_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 300 $NCU -k regex:"huff_lanes_kernel|lz_resolve_kernel" -c 2 -f -o gpurun_out/r02_c3_seg python tools/ab_bench.py --workloads c3 --streams 262144 --steps 0 --variants "v2:" > gpurun_out/r02_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
timeout 400 $NCU -k regex:"huff_lanes_kernel|lz_resolve_kernel" -c 2 -f -o gpurun_out/r02_c4_seg python tools/ab_bench.py --workloads c4 --streams 4096 --steps 0 --variants "v2:" > gpurun_out/r02_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
