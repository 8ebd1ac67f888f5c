#!/bin/bash
# Round 2, call AG (2 GPUs): both bench arms launched the way the driver launches them.
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r3g_ref_n2.json 2> gpurun_out/r3g_ref_n2.err
tail -c 600 gpurun_out/r3g_ref_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3g_n2.json 2> gpurun_out/r3g_n2.err
tail -3 gpurun_out/r3g_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3g_n2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "steps", "ms_per_step", "launch_mode", "eager_ms_per_step", "gpu_launches")})
print(d["e2e"]); print(d["clocks"]); print(d["roofline"]["frac"])
PY
