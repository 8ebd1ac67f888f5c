#!/bin/bash
# Round 2, call AC: why the bench's e2e leg (9.6 M) is below tools/e2e_ab.py (13 M) - both on the same box.
set -x
nproc
python tools/e2e_ab.py --threads 8,12,16 --steps 20 2>&1 | cut -c1-200
for th in default 12 8; do
  if [ "$th" = default ]; then unset TDE_HOST_THREADS; else export TDE_HOST_THREADS=$th; fi
  python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-other-configs --policy random 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench threads $th', d['value'], d['e2e']['by_mode'])"
done
unset TDE_HOST_THREADS
OMP_NUM_THREADS=1 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-other-configs --policy random 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench OMP_NUM_THREADS=1', d['value'], d['e2e']['by_mode'])"
