#!/bin/bash
# step-only and bench-workload timings for the tick kernel variants (FFMP_TICK_PIPE=0/1), optional library override
for pipe in 0 1; do
  echo "== pipe=$pipe"
  FFMP_TICK_PIPE=$pipe timeout 200 python tools/kbench.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: round(d[k], 2) for k in ('step_only_us', 'bench_us_per_step')})
    elif 'rror' in l: print(l.strip()[:200])"
done
