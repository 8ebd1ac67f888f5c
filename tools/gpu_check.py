"""Developer diagnostic: CUDA engine vs CPU oracle on a few configurations (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200.engine import Engine
from torchdriveenv_b200._capi import default_config
from oracle import oracle as O

def compare(name, ss, E, A, steps=60, seed=3, **cfg):
    eng = Engine(ss, E, A, **cfg)
    ocfg = default_config(num_envs=E, max_agents=A, **cfg)
    orc = O.OracleEnvSet(ocfg, eng.packed)
    eng.reset(seed=seed); orc.reset(seed=seed)
    torch.cuda.synchronize()
    st = eng.get_state().cpu().numpy()
    print(f"[{name}] reset state equal:", np.array_equal(st, orc.state), "vars equal:", np.array_equal(eng.get_env_vars().cpu().numpy(), orc.env_vars))
    obs0 = eng.render().cpu().numpy(); oobs0 = orc.render()
    print(f"[{name}] reset obs identical frac:", (obs0 == oobs0).mean())
    rng = np.random.default_rng(seed)
    bad = {}
    for k in range(steps):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, r, te, tr, info = eng.step(torch.from_numpy(a).cuda())
        oobs, orr, ote, otr, oinfo = orc.step(a)
        torch.cuda.synchronize()
        res = dict(state=np.array_equal(eng.get_state().cpu().numpy(), orc.state),
                   infr=np.array_equal(eng.get_infractions().cpu().numpy(), orc.infractions),
                   vars=np.array_equal(eng.get_env_vars().cpu().numpy(), orc.env_vars),
                   reward=np.array_equal(r.cpu().numpy(), orr), term=np.array_equal(te.cpu().numpy(), ote),
                   trunc=np.array_equal(tr.cpu().numpy(), otr), info=np.array_equal(info.cpu().numpy(), oinfo),
                   obs=float((obs.cpu().numpy() == oobs).mean()))
        for key, v in res.items():
            if (key == 'obs' and v < 1.0) or (key != 'obs' and not v):
                bad.setdefault(key, []).append((k, v))
    print(f"[{name}] mismatches over {steps} steps:", {k: (len(v), v[:3]) for k, v in bad.items()} or "none")
    if 'state' in bad:
        d = np.abs(eng.get_state().cpu().numpy() - orc.state); print("   max state abs diff", d.max())
    if 'infr' in bad:
        gi = eng.get_infractions().cpu().numpy(); d = np.abs(gi - orc.infractions); idx = np.unravel_index(d.argmax(), d.shape)
        print("   infr max diff", d.max(), idx, gi[idx], orc.infractions[idx])
    print(f"[{name}] stats gpu", eng.episode_stats()[:9], "\n           orc", orc.stats[:9])
    return eng, orc

compare("three_way E=4", S.three_way(), 4, 9)
compare("roundabout E=64 A=16", S.roundabout(16), 64, 16, auto_reset=1)
compare("traffic_lights E=256 A=32", S.traffic_lights(32), 256, 32, auto_reset=1)
compare("mix E=64 A=40", S.validation_mix(40), 64, 44, auto_reset=1, randomize_ego_attributes=1, steps=40)

# stateless kernels
st, at = S.scatter_boxes(512, 64, size=80.0, seed=1, present_p=0.9)
patch = S.ScenarioSet([S.scatter_patch(80.0, 10.0)], [S.make_scenario(0, [[5, 5], [50, 5]], 0, 0, "p")])
eng = Engine(patch, 4, 1)
g = eng.collision_boxes(torch.from_numpy(st), torch.from_numpy(at)).cpu().numpy()
o = O.collision_boxes(st, at)
print("collision_boxes equal:", np.array_equal(g, o), "colliding frac", (o > 0).mean())
g = eng.offroad_boxes(0, torch.from_numpy(st), torch.from_numpy(at)).cpu().numpy()
o = O.offroad_boxes(patch.maps[0].road_tris, 0.5, st, at)
print("offroad_boxes equal:", np.array_equal(g, o), "max diff", np.abs(g - o).max(), "offroad frac", (o > 0).mean())
# points far off the grid
st2 = st.copy(); st2[..., 0] += 300
g = eng.offroad_boxes(0, torch.from_numpy(st2), torch.from_numpy(at)).cpu().numpy()
o = O.offroad_boxes(patch.maps[0].road_tris, 0.5, st2, at)
print("offroad_boxes (off-grid) equal:", np.array_equal(g, o))

# quick timing, config C3
E, A = 16384, 32
eng = Engine(S.traffic_lights(32), E, A, auto_reset=1)
eng.reset(seed=0)
print("map info", eng.map_info(0))
act = torch.stack([torch.rand(E, device='cuda') * 2 - 1, torch.rand(E, device='cuda') * 0.6 - 0.3], 1)
for name, render in (("full step", True), ("no render", False)):
    for _ in range(5): eng.step(act, render=render)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(50): eng.step(act, render=render)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 50
    print(f"C3 {name}: {ms:.3f} ms/step -> {E / ms * 1e3 / 1e6:.2f} M env-steps/s")
ev0.record()
for _ in range(50): eng.render()
ev1.record(); torch.cuda.synchronize()
print(f"C3 render only: {ev0.elapsed_time(ev1) / 50:.3f} ms")
print("stats", eng.episode_stats()[:9])
