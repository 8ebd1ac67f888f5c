"""Would physics (latency-bound) and render (issue-bound) gain from sharing the SMs?  Two engines on two
streams: A steps physics only, B renders only.  Compare back-to-back at full grids with concurrent at
capped grids (TDE_PHYS_BLOCKS_CAP / TDE_RENDER_BLOCKS_CAP are read at tde_create)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200.engine import Engine
E, A = 16384, 32
rng = np.random.default_rng(0)
acts = torch.from_numpy(np.stack([rng.uniform(-1, 1, (64, E)), rng.uniform(-0.3, 0.3, (64, E))], -1).astype(np.float32)).cuda()
def make(pc, rc):
    for k, v in (("TDE_PHYS_BLOCKS_CAP", pc), ("TDE_RENDER_BLOCKS_CAP", rc)):
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)
    a = Engine(S.traffic_lights(A), E, A, device="cuda:0", auto_reset=1); a.reset(seed=0)
    b = Engine(S.traffic_lights(A), E, A, device="cuda:0", auto_reset=1); b.reset(seed=1)
    for k in range(20):
        a.step(acts[k % 64]); b.step(acts[(k + 7) % 64])
    return a, b
def run(a, b, concurrent, n=50):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if concurrent:
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        for k in range(n):
            with torch.cuda.stream(s1): a.step(acts[k % 64], render=False)
            with torch.cuda.stream(s2): b.render(out=b.obs)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    else:
        for k in range(n):
            a.step(acts[k % 64], render=False); b.render(out=b.obs)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
a, b = make(None, None)
print(f"full grids: back to back {run(a, b, False):.1f} us   two streams {run(a, b, True):.1f} us")
a.close(); b.close()
for pc, rc in ((4, 4), (3, 5), (2, 6), (5, 3)):
    a, b = make(pc, rc)
    print(f"caps physics {pc} render {rc} blocks/SM: back to back {run(a, b, False):.1f} us   two streams {run(a, b, True):.1f} us")
    a.close(); b.close()
