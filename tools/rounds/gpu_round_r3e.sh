#!/bin/bash
# Round 2, call AE: config C2 replayed from CUDA graphs against one Python call per step.
set -x
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -k "capture_steps" 2>&1 | tail -2
for g in on off; do
  python bench.py --workload c2 --steps 3000 --warmup 100 --no-cpu-baseline --cuda-graph $g 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graph $g', d['value'], d['ms_per_step'], d['launch_mode'], d['gpu_launches'], d['roofline']['frac'], d['clocks'])"
done
python bench.py --workload c3 --steps 500 --warmup 20 --no-cpu-baseline --no-other-configs --policy random --cuda-graph on 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3 graph on', d['value'], d['ms_per_step'])"
