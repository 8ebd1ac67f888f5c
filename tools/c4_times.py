"""C4 micro-benchmark: 65,536 envs x 64 agents, all-pairs SAT + offroad (CUDA events, warm)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200.engine import Engine
E, A = 65536, 64
size = float(sys.argv[1]) if len(sys.argv) > 1 else 200.0
st, at = S.scatter_boxes(E, A, size=size, seed=12)
patch = S.scatter_patch(size, 10.0)
eng = Engine(S.ScenarioSet([patch], [S.make_scenario(0, [[5, 5], [50, 5]], 0, 0, "p")]), 1, 1, device="cuda:0")
st_d, at_d = torch.from_numpy(st).cuda(), torch.from_numpy(at).cuda()
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
tc = timed(lambda: eng.collision_boxes(st_d, at_d))
to = timed(lambda: eng.offroad_boxes(0, st_d, at_d))
col = eng.collision_boxes(st_d, at_d)
print(f"C4 size={size}: collision {tc:.1f} us  offroad {to:.1f} us; algorithmic 121.6 MB -> {121.6e6 / ((tc + to) * 1e-6) / 1e9:.0f} GB/s; pairs overlapping {float(col.sum()) / 2 / (E * A * (A - 1) / 2) * 100:.2f}%  map {eng.map_info(0)}")
