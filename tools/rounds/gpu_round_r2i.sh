#!/bin/bash
# Round 2, call I: rollout collector with CUDA-graph replay (test + C5 bench), kernel times of the C5 step.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_env_api.py -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_i.txt
python bench.py --workload c5 --steps 512 --warmup 96 --no-cpu-baseline | tee gpurun_out/bench_c5.json
python - <<'PY' 2>&1 | tee gpurun_out/c5_kernel_times.txt
import sys, numpy as np, torch
sys.path.insert(0, '.')
from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200.engine import Engine
E, A, n = 8192, 8, 3
eng = Engine(S.training_mix(100, A), E, A, device="cuda:0", auto_reset=1)
eng.reset(seed=0)
rng = np.random.default_rng(0)
acts = torch.from_numpy(np.stack([rng.uniform(-1, 1, (64, E)), rng.uniform(-0.3, 0.3, (64, E))], -1).astype(np.float32)).cuda()
for k in range(20): eng.step(acts[k % 64])
def timed(fn, nrep=100):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(nrep): fn(k)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / nrep * 1e3
buf = torch.zeros((33, E, 9, 64, 64), dtype=torch.uint8, device="cuda")
obs = eng.render()
print("C5 shard: physics %.1f us, render %.1f us, step %.1f us, scatter step %.1f us, shift step %.1f us" % (
    timed(lambda k: eng.step(acts[k % 64], render=False)), timed(lambda k: eng.render(out=obs)), timed(lambda k: eng.step(acts[k % 64])),
    timed(lambda k: eng.step_rollout_scatter(acts[k % 64], buf, k % 32, n)), timed(lambda k: eng.step_rollout(acts[k % 64], buf[k % 32], buf[k % 32 + 1], n))))
PY
