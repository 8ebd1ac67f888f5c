"""Per-env start/end timestamps of one C3 step (debug build: tools/build_variant.sh trace -DTDE_TRACE [...]):
TDE_B200_LIB=variants/lib_trace.so python tools/trace_envs.py [E] [A]"""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from torchdriveenv_b200 import scenarios as S, _capi
from torchdriveenv_b200.engine import Engine
E = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
A = int(sys.argv[2]) if len(sys.argv) > 2 else 32
eng = Engine(S.traffic_lights(A), E, A, device="cuda:0", auto_reset=1)
eng.reset(seed=0)
rng = np.random.default_rng(0)
acts = torch.from_numpy(np.stack([rng.uniform(-1, 1, (64, E)), rng.uniform(-0.3, 0.3, (64, E))], -1).astype(np.float32)).cuda()
for k in range(20):
    eng.step(acts[k % 64])
trace = torch.zeros((E, 8), dtype=torch.int64, device="cuda")
lib = _capi.load_library()
lib.tde_debug_set_trace.argtypes = [ctypes.c_void_p]
assert lib.tde_debug_set_trace(trace.data_ptr()) == 0
torch.cuda.synchronize()
eng.step(acts[21])
torch.cuda.synchronize()
t = trace.cpu().numpy().astype(np.float64)
cnt = t[:, 4:7].copy()
t = t[:, :4]
t0 = t[t > 0].min()
t = (t - t0) / 1e3
for name, a, b in (("physics", 2, 3), ("render", 0, 1)):
    st, en = t[:, a], t[:, b]
    d = en - st
    print(f"{name}: first start {st.min():.1f} us, last start {st.max():.1f}, last end {en.max():.1f}; per-env duration mean {d.mean():.1f} "
          f"p10 {np.percentile(d,10):.1f} p50 {np.percentile(d,50):.1f} p90 {np.percentile(d,90):.1f} max {d.max():.1f}")
    # duration by start-time bucket, and number of envs in flight over time
    edges = np.linspace(st.min(), en.max(), 13)
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = (st >= lo) & (st < hi)
        mid = 0.5 * (lo + hi)
        inflight = int(((st <= mid) & (en > mid)).sum())
        print(f"   t=[{lo:6.1f},{hi:6.1f}) started {int(sel.sum()):6d}  mean duration {d[sel].mean() if sel.any() else 0:6.1f} us   in flight at mid {inflight}")
d = t[:, 1] - t[:, 0]
late = t[:, 0] > np.percentile(t[:, 0], 35)   # skip the first wave (everything starts at once there)
for name, x in (("static queued", cnt[:, 0]), ("dynamic items", cnt[:, 1]), ("queued total", cnt[:, 2])):
    print(f"{name}: mean {x.mean():.1f} p10 {np.percentile(x,10):.0f} p50 {np.percentile(x,50):.0f} p90 {np.percentile(x,90):.0f} max {x.max():.0f}; corr with render duration (after the first wave) {np.corrcoef(x[late], d[late])[0,1]:.3f}")
A_ = np.stack([cnt[late, 0], cnt[late, 1], np.ones(late.sum())], 1)
coef, *_ = np.linalg.lstsq(A_, d[late], rcond=None)
print("least squares: duration ~ %.3f*static + %.3f*dynamic + %.2f us; residual std %.2f" % (*coef, (A_ @ coef - d[late]).std()))
