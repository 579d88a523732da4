#!/bin/bash
# one gpurun call: the queue mode (pass 2 beside pass 1) at growing batch sizes, each under its own timeout
mkdir -p gpurun_out
TAG=${1:-q}
SIZES=${2:-"4096 30000 65536"}
VARIANTS=${3:-"q4:SFB200_QUEUE=1;q2:SFB200_QUEUE=1,SFB200_QUEUE_CTAS=2;q1:SFB200_QUEUE=1,SFB200_QUEUE_CTAS=1;base:"}
WL=${4:-c2}
for N in $SIZES; do
  timeout 150 python tools/ab_bench.py --workloads $WL --streams $N --unique 512 --steps 5 \
    --variants "$VARIANTS" --out gpurun_out/r02_ab_$TAG.jsonl 2>> gpurun_out/r02_ab_$TAG.err | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['workload'], d['n'], d['variant'], 'ok' if d['ok'] else 'WRONG', 'p1 %.2f p2 %.2f step %.2f ms  %.1f GB/s' % (d['pass1_ms'], d['pass2_ms'], d['step_ms'], d['gbs']), d.get('queue'))
"
  echo "N=$N rc=$?"
done
tail -5 gpurun_out/r02_ab_$TAG.err
