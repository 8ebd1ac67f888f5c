"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200.engine import Engine
for (ss, E, A) in ((S.traffic_lights(32), 96, 32), (S.validation_mix(44), 40, 48)):
    eng = Engine(ss, E, A, device="cuda:0", auto_reset=1, max_environment_steps=6)
    eng.reset(seed=3)
    rng = np.random.default_rng(3)
    stack = torch.zeros((E, 9, 64, 64), dtype=torch.uint8, device="cuda")
    eng.render_stacked(stack, 3)
    for k in range(8):
        a = torch.from_numpy(np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)).cuda()
        eng.step(a)
        eng.step_stacked(a, stack, 3)
    st, at = S.scatter_boxes(64, 64, size=100.0, seed=1)
    eng.collision_boxes(torch.from_numpy(st).cuda(), torch.from_numpy(at).cuda())
    eng.offroad_boxes(0, torch.from_numpy(st).cuda(), torch.from_numpy(at).cuda())
    torch.cuda.synchronize()
print("sanitize run done")
