#!/bin/bash
# one gpurun call: GPU parity suite, then an A/B of pass-2 kernels on C2/C3/C4
set -x
mkdir -p gpurun_out
TAG=${1:-x}
VARIANTS=${2:-"v2:;v1:SFB200_LZ_V1=1"}
WL=${3:-c2,c4,c3}
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_$TAG.log 2>&1; echo "pytest rc=$?"; fi
[ -z "$SKIP_TESTS" ] && tail -5 gpurun_out/r02_pytest_$TAG.log
python tools/ab_bench.py --workloads $WL --variants "$VARIANTS" --steps 5 --unique 2048 --out gpurun_out/r02_ab_$TAG.jsonl 2> gpurun_out/r02_ab_$TAG.err | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['workload'], d['variant'], 'ok' if d['ok'] else 'WRONG', 'p1 %.2f p2 %.2f step %.2f ms  %.1f GB/s' % (d['pass1_ms'], d['pass2_ms'], d['step_ms'], d['gbs']), 'lz regs', d['launch'].get('lz_regs_per_thread'), 'ctas', d['launch'].get('lz_ctas_per_sm'))
"
tail -3 gpurun_out/r02_ab_$TAG.err
