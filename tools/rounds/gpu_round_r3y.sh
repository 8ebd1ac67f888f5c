#!/bin/bash
# Round 2, call AY (4 GPUs): both bench arms of the final build, launched the way the driver launches them.
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r3y_n4.json 2> gpurun_out/r3y_n4.err
tail -2 gpurun_out/r3y_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > gpurun_out/r3y_ref_n4.json 2> gpurun_out/r3y_ref_n4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3y_n4.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "steps", "ms_per_step", "launch_mode", "eager_ms_per_step", "gpu_launches")})
print(d["e2e"]); print(d["clocks"]); print(d["roofline"]["frac"])
r = json.loads(open("gpurun_out/r3y_ref_n4.json").read().strip().splitlines()[-1])
print({k: r[k] for k in ("value", "n_gpus", "ms_per_step")}, r["cpu_baseline"]["cores"])
PY
