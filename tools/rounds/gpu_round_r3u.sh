#!/bin/bash
# Round 2, call AU: episode statistics added per CTA (shared-memory atomics, then one set of global atomics) instead of per warp.
set -x
for v in base ctastats base ctastats; do
  if [ "$v" = base ]; then unset TDE_B200_LIB; else export TDE_B200_LIB=$PWD/variants/lib_$v.so; fi
  python bench.py --workload c2 --steps 3000 --warmup 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v c2', d['value'], d['ms_per_step'], d['eager_ms_per_step'])"
done
tools/ab_checked.sh base ctastats
