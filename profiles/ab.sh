#!/bin/bash
# A/B of gather variants on one B200: prints kernel ms per launch for each setting (bench.py --no-e2e --no-cpu)
EDGE=${1:-216}
shift
for v in "$@"; do
  env $(echo $v | tr ',' ' ') python bench.py --edge $EDGE --steps 2 --warmup 2 --no-e2e --no-cpu 2>gpurun_out/ab_err.log | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        j=json.loads(ln); k=j['roofline']['kernel_ms_per_launch']
        print('$v', 'ms/step', j['ms_per_step'], {a:k[a] for a in k if a.startswith('gather')})
"
done
