#!/bin/bash
# Round 2, call AP: all GPU tests with the auto-staged small batches, smoke, the C2 / C1 numbers.
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --workload c2 --steps 3000 --warmup 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['ms_per_step'], d['eager_ms_per_step'])"
for e in 256 512 1184 1185; do python tools/kernel_times.py $e 16 | head -1; TDE_PHYS_STAGE=0 python tools/kernel_times.py $e 16 | head -1; done
