#!/bin/bash
# Round 2, call Q: C5 rollout collection with the ring of frames against the scatter store.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "ring_of_frames or rollout" 2>&1 | tail -3
for mode in ring scatter ring; do
  timeout 600 python bench.py --workload c5 --steps 512 --warmup 64 --no-cpu-baseline --c5-frame-copy $mode > gpurun_out/r2q_c5_$mode.json 2> gpurun_out/r2q_c5_$mode.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2q_c5_$mode.json").read().strip().splitlines()[-1])
print("$mode", d["value"], d["ms_per_step"], d["clocks"], d["e2e"]["value"], d["roofline"]["frac"])
PY
done
