import numpy as np, torch, sys
sys.path.insert(0, '.')
from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200.engine import Engine
E, A, steps = 16384, 32, 12
ss = S.traffic_lights(A)
rng = np.random.default_rng(10)
acts = np.stack([rng.uniform(-1, 1, (steps, E)), rng.uniform(-0.3, 0.3, (steps, E))], -1).astype(np.float32)
eng = Engine(ss, E, A, device="cuda:0", auto_reset=1)
eng.reset(seed=10)
for k in range(steps):
    eng.step(torch.from_numpy(acts[k]).cuda())
st = eng.get_state().cpu().numpy()
print("finite", np.isfinite(st).all(), "psi min", st[..., 2].min(), "psi max", st[..., 2].max(), np.pi)
bad = ~((st[..., 2] >= -np.pi - 1e-6) & (st[..., 2] < np.pi + 1e-6))
print("bad count", bad.sum(), np.argwhere(bad)[:10], st[bad][:10])
