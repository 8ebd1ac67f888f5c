#!/bin/bash
# Round 2, call AF: render-less steps launched with programmatic stream serialization (C2), graphs on every step workload.
set -x
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for pdl in 1 0; do for g in on off; do
  TDE_PDL_PHYSICS=$pdl python bench.py --workload c2 --steps 3000 --warmup 100 --no-cpu-baseline --cuda-graph $g 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl $pdl graph $g', d['value'], d['ms_per_step'], d['eager_ms_per_step'], d['gpu_launches'])"
done; done
python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --no-other-configs --policy random 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'], d['eager_ms_per_step'], d['launch_mode'], d['roofline']['frac'], d['gpu_launches'])"
