#!/bin/bash
# Round 2, call AO: config C2 with the map tables staged into shared memory by bulk-async copies (cold L1 at every launch).
set -x
for st in 0 1; do
  TDE_PHYS_STAGE=$st python bench.py --workload c2 --steps 3000 --warmup 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stage $st', d['value'], d['ms_per_step'], d['eager_ms_per_step'])"
done
for st in 0 1; do TDE_PHYS_STAGE=$st python tools/kernel_times.py 1024 16 | head -1; TDE_PHYS_STAGE=$st python tools/kernel_times.py 2048 16 | head -1; done
