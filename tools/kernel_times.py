"""Per-kernel device times of the C3 step (CUDA events, warm): python tools/kernel_times.py [E] [A]"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200.engine import Engine
E = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
A = int(sys.argv[2]) if len(sys.argv) > 2 else 32
eng = Engine(S.traffic_lights(A), E, A, device="cuda:0", auto_reset=1)
eng.reset(seed=0)
rng = np.random.default_rng(0)
acts = torch.from_numpy(np.stack([rng.uniform(-1, 1, (64, E)), rng.uniform(-0.3, 0.3, (64, E))], -1).astype(np.float32)).cuda()
for k in range(20):
    eng.step(acts[k % 64])
obs = eng.render()
def timed(fn, n=50):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n):
        fn(k)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t_phys = timed(lambda k: eng.step(acts[k % 64], render=False))
t_rend = timed(lambda k: eng.render(out=obs))
t_both = timed(lambda k: eng.step(acts[k % 64]))
stack = torch.zeros((E, 9, 64, 64), dtype=torch.uint8, device="cuda")
t_stack = timed(lambda k: eng.step_stacked(acts[k % 64], stack, 3))
ring = eng.new_stack_ring(3)
eng.render_stacked_ring(ring, 3)
t_ring = timed(lambda k: eng.step_stacked_ring(acts[k % 64], ring, k + 1, 3))
print(f"E={E} A={A} physics {t_phys:.1f} us  render {t_rend:.1f} us  step {t_both:.1f} us  ({E / t_both:.1f} M env-steps/s)  step with fused 3-frame stack {t_stack:.1f} us  with the ring of stacks {t_ring:.1f} us")
print("map", eng.map_info(0))
