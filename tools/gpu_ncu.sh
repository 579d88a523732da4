#!/bin/bash
# one gpurun call: ncu --set full of one kernel of one workload (after a plain run has exited 0)
# usage: gpu_ncu.sh TAG WORKLOAD STREAMS KERNEL_REGEX [ENV=VAL,...]
set -x
mkdir -p gpurun_out
TAG=$1; WL=$2; N=$3; KRE=$4; ENVS=${5:-}
python tools/ab_bench.py --workloads $WL --streams $N --variants "p:$ENVS" --steps 1 --unique 1024 > gpurun_out/prof_$TAG.plain.log 2>&1 || { tail -5 gpurun_out/prof_$TAG.plain.log; exit 1; }
tail -1 gpurun_out/prof_$TAG.plain.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:$KRE --launch-skip ${SKIP:-2} --launch-count ${COUNT:-1} -f -o gpurun_out/prof_$TAG \
  python tools/ab_bench.py --workloads $WL --streams $N --variants "p:$ENVS" --steps 1 --unique 1024 > gpurun_out/prof_$TAG.ncu.log 2>&1
tail -3 gpurun_out/prof_$TAG.ncu.log
ls -la gpurun_out/prof_$TAG.ncu-rep
