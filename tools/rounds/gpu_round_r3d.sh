#!/bin/bash
# Round 2, call AD: where a lone warp waits in the physics kernel at config C2 (1,024 envs x 16 agents): ncu source page.
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:tde_physics_kernel' -s 200 -c 1 -o gpurun_out/prof_c2 -f \
    python bench.py --workload c2 --steps 300 --warmup 50 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_c2.log 2>&1
tail -1 gpurun_out/ncu_c2.log | cut -c1-200
python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200.engine import Engine
from torchdriveenv_b200._capi import PH_ALL
E, A = 1024, 16
eng = Engine(S.roundabout(A), E, A, device="cuda:0", auto_reset=1)
eng.reset(seed=0)
acts = torch.zeros((E, 2), device="cuda")
def timed(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for name, ph in (("kinematics", 1), ("infractions", 2), ("kin+infr", 3), ("reward", 4), ("all no render", 7)):
    print(name, round(timed(lambda: eng.step(acts, render=False, phases=ph | 0)), 2), "us")
PY
