"""GPU parity tests added in round 2: the holes the round-1 review listed (VERDICT "Close the test holes").

(a) collision pile-ups that overflow the compacted pair list (TDE_PAIR_CAP), stateless and inside the step;
(b) the whole C3 batch (16,384 envs x 32 agents) against the oracle for two steps, every env, bit for bit;
(c) C5 at its bench size (8,192 envs x 8 agents, 100-map training mix) on env sub-ranges;
(d) a batch under a pursuit policy that survives > 100 steps: junctions, red lights and late waypoints rendered and scored;
(e) NaN / inf / huge actions stay inside their env;
(f) tde_clone: the copy steps bit for bit like the original;
(g) the physics launch with the map tables staged in shared memory (bulk-async copies) against the default launch."""
import numpy as np
import pytest
import torch

from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200._capi import default_config

pytestmark = pytest.mark.gpu


def _engine(ss, E, A, **cfg):
    from torchdriveenv_b200.engine import Engine
    return Engine(ss, E, A, device="cuda:0", **cfg)


def pursuit_actions(state, env_vars, waypoints, rng, v_want=8.0):
    """Pure pursuit of the current target waypoint (the driver of tests/golden/make_reference_golden.py, vectorised)."""
    x, y, psi, v = (state[:, 0, k] for k in range(4))
    tgt = np.clip(env_vars[:, 2], 0, len(waypoints) - 1)
    w = waypoints[tgt]
    err = np.arctan2(w[:, 1] - y, w[:, 0] - x) - psi
    err = (err + np.pi) % (2 * np.pi) - np.pi
    acc = np.clip(0.8 * (v_want - v), -1, 1)
    steer = np.clip(0.35 * err + rng.normal(0, 0.02, len(x)), -0.3, 0.3)
    return np.stack([acc, steer], 1).astype(np.float32)


def test_collision_pile_up_overflows_the_pair_list(oracle):
    """64 agents in a 15 m square: hundreds of candidate pairs per env, far beyond the 256 the compacted list holds, so the
    per-lane fallback of sat_counts runs - stateless (tde_collision_boxes) and inside the step."""
    st, at = S.scatter_boxes(96, 64, size=15.0, seed=4)
    want = oracle.collision_boxes(st, at)
    assert want.max() >= 20 and (want.sum(1) / 2).min() > 256         # more overlapping pairs alone than the list holds
    patch = S.scatter_patch(60.0, 10.0)
    sc = S.make_scenario(0, [[5, 5], [50, 5]], 63, 0, "pile")
    eng = _engine(S.ScenarioSet([patch], [sc]), 96, 64)
    got = eng.collision_boxes(torch.from_numpy(st).cuda(), torch.from_numpy(at).cuda()).cpu().numpy()
    assert np.array_equal(got, want)
    # the same boxes as the env state: the in-step path
    orc = oracle.OracleEnvSet(default_config(num_envs=96, max_agents=64), eng.packed)
    eng.reset(seed=1); orc.reset(seed=1)
    st2 = st.copy(); st2[..., 0] += 20.0; st2[..., 1] += 20.0
    eng.set_state(torch.from_numpy(st2)); orc.state[...] = st2
    eng.set_attributes(torch.from_numpy(at)); orc.attr[...] = at
    assert np.array_equal(eng.compute_infractions().cpu().numpy(), orc.compute_infractions())
    assert orc.infractions[..., 0].max() >= 20


def test_whole_c3_batch_two_steps(oracle):
    """BASELINE config C3 at its full size: every one of the 16,384 envs against the oracle, reset + two steps."""
    E, A = 16384, 32
    eng = _engine(S.traffic_lights(A), E, A, auto_reset=1)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1), eng.packed)
    eng.reset(seed=21); orc.reset(seed=21)
    assert np.array_equal(eng.get_state().cpu().numpy(), orc.state)
    assert np.array_equal(eng.render().cpu().numpy(), orc.render())
    rng = np.random.default_rng(21)
    for k in range(2):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, r, te, tr, info = eng.step(torch.from_numpy(a).cuda())
        oobs, orr, ote, otr, oinfo = orc.step(a)
        assert np.array_equal(eng.get_state().cpu().numpy(), orc.state), f"step {k}: state"
        assert np.array_equal(eng.get_infractions().cpu().numpy(), orc.infractions), f"step {k}: infractions"
        assert np.array_equal(info.cpu().numpy(), oinfo) and np.array_equal(r.cpu().numpy(), orr), f"step {k}: info / reward"
        assert np.array_equal(te.cpu().numpy(), ote) and np.array_equal(tr.cpu().numpy(), otr), f"step {k}: flags"
        bad = (obs.cpu().numpy() != oobs).reshape(E, -1).any(1)
        assert not bad.any(), f"step {k}: {int(bad.sum())} envs render differently"


def test_c5_bench_size_on_env_sub_ranges(oracle):
    """BASELINE config C5's per-GPU shard (8,192 envs x 8 agents, 100 training maps - the physics reads the tables from
    global memory: they do not fit in shared memory): three windows of 96 envs against the oracle through the stacked step."""
    E, A, n = 8192, 8, 3
    ss = S.training_mix(100, A)
    eng = _engine(ss, E, A, auto_reset=1)
    eng.reset(seed=9)
    stack = torch.zeros((E, 3 * n, 64, 64), dtype=torch.uint8, device="cuda")
    eng.render_stacked(stack, n)
    windows = [(0, 96), (4000, 4096), (E - 96, E)]
    orcs = []
    for lo, hi in windows:
        o = oracle.OracleEnvSet(default_config(num_envs=hi - lo, max_agents=A, auto_reset=1, env_index_offset=lo), eng.packed)
        o.reset(seed=9)
        assert np.array_equal(eng.get_state()[lo:hi].cpu().numpy(), o.state)
        orcs.append(o)
    rng = np.random.default_rng(9)
    for k in range(12):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        eng.step_stacked(torch.from_numpy(a).cuda(), stack, n)
        st = eng.get_state().cpu().numpy(); info = eng.info.cpu().numpy(); newest = stack[:, -3:].cpu().numpy()
        for (lo, hi), o in zip(windows, orcs):
            oobs, orr, ote, otr, oinfo = o.step(a[lo:hi])
            assert np.array_equal(st[lo:hi], o.state), f"step {k} window {lo}: state"
            assert np.array_equal(info[lo:hi], oinfo), f"step {k} window {lo}: info"
            assert np.array_equal(newest[lo:hi], oobs), f"step {k} window {lo}: newest frame"
    assert len(set(eng.get_env_vars()[:, 0].cpu().numpy().tolist())) > 90      # the shard draws (nearly) all 100 scenarios


def test_pursuit_policy_survives_and_is_scored(oracle):
    """Envs driven by pure pursuit of their target waypoint: most survive > 100 steps, pass the junctions, meet red lights
    and collect waypoints - the expensive end of the render / scoring distribution, bit for bit against the oracle."""
    E, A = 128, 16
    ss = S.traffic_lights(A)
    cfg = dict(auto_reset=1, terminated_at_infraction=0, distance_cutoff=0.25)
    eng = _engine(ss, E, A, **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, **cfg), eng.packed)
    eng.reset(seed=31); orc.reset(seed=31)
    wps = np.asarray(ss.scenarios[0].waypoints, np.float64)
    rng = np.random.default_rng(31)
    reached, red, max_len = 0, 0, 0
    for k in range(150):
        a = pursuit_actions(orc.state.astype(np.float64), orc.env_vars, wps, rng)
        obs, r, te, tr, info = eng.step(torch.from_numpy(a).cuda())
        oobs, orr, ote, otr, oinfo = orc.step(a)
        assert np.array_equal(eng.get_state().cpu().numpy(), orc.state), f"step {k}: state"
        assert np.array_equal(info.cpu().numpy(), oinfo), f"step {k}: info"
        assert np.array_equal(obs.cpu().numpy(), oobs), f"step {k}: observation"
        assert np.array_equal(te.cpu().numpy(), ote) and np.array_equal(tr.cpu().numpy(), otr)
        reached = max(reached, int(oinfo[:, 4].max())); red += int((oinfo[:, 2] > 0).sum())
        max_len = max(max_len, int(oinfo[:, 11].max()))
    assert max_len >= 140 and reached >= 5 and red > 0, (max_len, reached, red)
    np.testing.assert_allclose(eng.episode_stats(), orc.stats, rtol=1e-9)


def test_non_finite_actions_stay_inside_their_env(oracle):
    """NaN, +-inf and huge actions in some envs: those envs may become garbage (as they do in the oracle, bit for bit where
    the result is defined), the others must be untouched."""
    E, A = 64, 16
    eng = _engine(S.traffic_lights(A), E, A, auto_reset=1)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1), eng.packed)
    eng.reset(seed=5); orc.reset(seed=5)
    rng = np.random.default_rng(5)
    bad_envs = np.array([3, 17, 18, 40, 63])
    good = np.ones(E, bool); good[bad_envs] = False
    poison = [np.nan, np.inf, -np.inf, 1e30, -1e38]
    for k in range(10):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        for j, e in enumerate(bad_envs):
            a[e, (j + k) % 2] = poison[(j + k) % len(poison)]
        obs, r, te, tr, info = eng.step(torch.from_numpy(a).cuda())
        oobs, orr, ote, otr, oinfo = orc.step(a)
        torch.cuda.synchronize()
        assert np.array_equal(eng.get_state().cpu().numpy()[good], orc.state[good]), f"step {k}"
        assert np.array_equal(info.cpu().numpy()[good], oinfo[good]) and np.array_equal(obs.cpu().numpy()[good], oobs[good]), f"step {k}"
        assert np.array_equal(te.cpu().numpy()[good], ote[good]) and np.array_equal(tr.cpu().numpy()[good], otr[good])
        # the poisoned envs agree too wherever the oracle's value is a number
        st, ost = eng.get_state().cpu().numpy()[~good], orc.state[~good]
        fin = np.isfinite(ost)
        assert np.array_equal(st[fin], ost[fin]) and np.array_equal(np.isnan(st), np.isnan(ost))


def test_clone_steps_like_the_original(oracle):
    """tde_clone (simulator.copy(), gym_env.py:110): device-to-device copy of every env, scenario tables shared."""
    E, A = 200, 16
    eng = _engine(S.validation_mix(10), E, A, auto_reset=1)
    eng.reset(seed=3)
    rng = np.random.default_rng(3)
    acts = [torch.from_numpy(np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)).cuda() for _ in range(12)]
    for a in acts[:4]:
        eng.step(a)
    twin = eng.clone()
    assert torch.equal(twin.get_state(), eng.get_state()) and torch.equal(twin.get_env_vars(), eng.get_env_vars())
    for a in acts[4:8]:
        o1, r1, t1, u1, i1 = eng.step(a)
        o2, r2, t2, u2, i2 = twin.step(a)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(i1, i2) and torch.equal(t1, t2) and torch.equal(u1, u2)
        assert torch.equal(twin.get_state(), eng.get_state())
    keep = eng.get_state().clone()
    for a in acts[8:]:
        twin.step(a)                                   # the copy moves on alone ...
    assert torch.equal(eng.get_state(), keep)          # ... and the original stays where it was
    eng.close()                                        # the shared tables outlive the first handle
    twin.step(acts[0])
    torch.cuda.synchronize()
    assert torch.cuda.current_device() == 0
    twin.close()


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("E,A", [(700, 32), (96, 64), (1, 9), (1500, 16)])
def test_staged_map_tables_match_the_default_launch(oracle, E, A, mode):
    """cfg.stage_map_tables = 1: the lane-mesh triangle records, stop lines, light schedule and the per-cell summary are
    copied into shared memory once per CTA (cp.async.bulk + mbarrier); 2: they are read through L1 (what a large batch does
    by default; small batches are staged by default).  Results equal the oracle's either way."""
    ss = S.three_way(6) if A == 9 else S.traffic_lights(A)
    cfg = dict(auto_reset=1, stage_map_tables=mode)
    eng = _engine(ss, E, A, **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1), eng.packed)
    eng.reset(seed=8); orc.reset(seed=8)
    rng = np.random.default_rng(8)
    for k in range(20):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, r, te, tr, info = eng.step(torch.from_numpy(a).cuda())
        oobs, orr, ote, otr, oinfo = orc.step(a)
        assert np.array_equal(eng.get_state().cpu().numpy(), orc.state), f"step {k}: state"
        assert np.array_equal(eng.get_infractions().cpu().numpy(), orc.infractions), f"step {k}: infractions"
        assert np.array_equal(info.cpu().numpy(), oinfo) and np.array_equal(obs.cpu().numpy(), oobs), f"step {k}"


def test_rollout_collector_cuda_graph_replay(oracle):
    """RolloutCollector(cuda_graph=True): the first rollout runs eagerly, the second is captured into one CUDA graph and
    replayed, the following ones only replay it.  Every buffer slot of five rollouts must hold what the eager collector
    (and VecFrameStack over the oracle's frames) produces."""
    from torchdriveenv_b200.engine import Engine
    from torchdriveenv_b200.rollout import RolloutCollector
    E, A, T, NS, R = 48, 6, 6, 3, 5
    ss = S.traffic_lights(A)
    cfg = dict(auto_reset=1, max_environment_steps=10)
    eng = Engine(ss, E, A, device="cuda:0", **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, **cfg), eng.packed)
    orc.reset(seed=19)
    col = RolloutCollector(eng, T, n_stack=NS, seed=19, cuda_graph=True)
    rng = np.random.default_rng(19)
    tape = torch.from_numpy(np.stack([rng.uniform(-1, 1, (R * T, E)), rng.uniform(-0.3, 0.3, (R * T, E))], -1).astype(np.float32)).cuda()
    cur = torch.zeros((T, E, 2), dtype=torch.float32, device="cuda")
    state = dict(r=0, t=0)

    def policy(obs, out=None):
        t = state["t"]; state["t"] += 1
        return out.copy_(cur[t])

    def before_rollout(n_steps):
        cur.copy_(tape[state["r"] * T:(state["r"] + 1) * T]); state["r"] += 1; state["t"] = 0

    policy.writes_into, policy.before_rollout = True, before_rollout
    frames, age = [orc.render()], np.zeros(E, np.int64)

    def want_stack():
        w = np.zeros((E, 3 * NS, 64, 64), np.uint8)
        for slot in range(NS):
            back = NS - 1 - slot
            if back < len(frames):
                have = age >= back
                w[have, 3 * slot:3 * slot + 3] = frames[-1 - back][have]
        return w

    for r in range(R):
        b = col.collect(policy)
        torch.cuda.synchronize()
        obs = b.observations.cpu().numpy()
        assert np.array_equal(obs[0], want_stack()), f"rollout {r}: slot 0"
        for t in range(T):
            a = tape[r * T + t].cpu().numpy()
            oobs, orr, ote, otr, oinfo = orc.step(a)
            d = (ote | otr).astype(bool)
            frames.append(oobs); age = np.where(d, 0, age + 1)
            assert np.array_equal(obs[t + 1], want_stack()), f"rollout {r} step {t}"
            assert np.array_equal(b.rewards[t].cpu().numpy(), orr) and np.array_equal(b.actions[t].cpu().numpy(), a)
            assert np.array_equal(b.episode_starts[t + 1].cpu().numpy().astype(bool), d)
    assert col._graph is not None and col.num_timesteps == R * T * E
    np.testing.assert_allclose(eng.episode_stats(), orc.stats, rtol=1e-9)


def test_compact_host_step_fills_the_same_bytes_as_the_rgb_host_step(oracle, monkeypatch):
    """tde_step_host by default: the frames cross PCIe as the 4-bit class image and host threads apply the palette; the
    caller's buffers must hold the same bytes as with cfg.host_obs_rgb = 1 (the planes cross PCIe) and as the oracle's, chunked (4,608 envs -> 8 chunks)
    and unchunked, and tde_render_classes must be the oracle's class image."""
    E, A = 4608, 16
    ss = S.validation_mix(12)
    cfg = dict(auto_reset=1)
    rgb = _engine(ss, E, A, host_obs_rgb=1, **cfg)
    cmp_ = _engine(ss, E, A, **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, **cfg), rgb.packed)
    for eng in (rgb, cmp_, orc):
        eng.reset(seed=12)
    assert np.array_equal(cmp_.render_classes().cpu().numpy(), orc.render_classes())
    rng = np.random.default_rng(3)
    for k in range(6):
        if k == 4:
            monkeypatch.setenv("TDE_HOST_CHUNKS", "1")
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        want = [np.array(x, copy=True) for x in rgb.step_host(a)]
        got = cmp_.step_host(a)
        ref = orc.step(a)
        for name, w, g, o in zip(("obs", "reward", "terminated", "truncated", "info"), want, got, ref):
            assert np.array_equal(w, g), f"step {k}: {name} differs between the RGB and the compact host step"
            assert np.array_equal(g, o), f"step {k}: {name} differs from the oracle"
    assert np.array_equal(cmp_.render_classes().cpu().numpy(), orc.render_classes())


def test_ring_of_frames_rollout_equals_the_scatter_rollout():
    """RolloutCollector(frame_copy="ring") keeps every frame once; what RolloutBuffer.stacked hands out must be the stacked
    observations the scatter collector stores, over several rollouts (carry-over, CUDA graph replay from the third on), for
    a policy that ignores the observation and for one that reads it at every step."""
    from torchdriveenv_b200.engine import Engine
    from torchdriveenv_b200.rollout import RolloutCollector, uniform_policy
    E, A, T, NS = 96, 6, 5, 3
    ss = S.traffic_lights(A)
    cfg = dict(auto_reset=1, max_environment_steps=9)

    def brightness_policy(obs):     # depends on every channel group of the stacked observation
        m = obs.view(obs.shape[0], -1).float().mean(1) / 255.0
        return torch.stack((torch.cos(40.0 * m), 0.3 * torch.sin(55.0 * m)), 1)

    for make_policy, graph in ((lambda: uniform_policy(seed=3), True), (lambda: brightness_policy, False)):
        cols = []
        for mode in ("scatter", "ring"):
            eng = Engine(ss, E, A, device="cuda:0", **cfg)
            cols.append(RolloutCollector(eng, T, n_stack=NS, seed=21, frame_copy=mode, cuda_graph=graph))
        pols = [make_policy(), make_policy()]
        n_done = 0
        for r in range(5):
            bs, br = cols[0].collect(pols[0]), cols[1].collect(pols[1])
            torch.cuda.synchronize()
            for t in range(T + 1):
                assert torch.equal(br.stacked(t), bs.observations[t]), f"rollout {r} observation {t}"
            for name in ("actions", "rewards", "terminated", "truncated", "episode_starts"):
                assert torch.equal(getattr(br, name), getattr(bs, name)), name
            n_done += int(bs.episode_starts[1:].sum())
        assert n_done > 20
        assert torch.equal(cols[1].last_observation(), cols[0].last_observation())


def test_capture_steps_replays_the_eager_steps(oracle):
    """Engine.capture_steps: a CUDA graph of G consecutive steps leaves the envs exactly where G eager steps (and the
    oracle) leave them, replay after replay."""
    E, A, G = 200, 16, 7
    ss = S.roundabout(A)
    cfg = dict(auto_reset=1)
    eager, graphed = _engine(ss, E, A, **cfg), _engine(ss, E, A, **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, **cfg), eager.packed)
    for eng in (eager, graphed, orc):
        eng.reset(seed=6)
    rng = np.random.default_rng(2)
    seq = torch.zeros((G, E, 2), dtype=torch.float32, device="cuda:0")
    graph = graphed.capture_steps(seq, render=True)
    for r in range(4):
        a = np.stack([rng.uniform(-1, 1, (G, E)), rng.uniform(-0.3, 0.3, (G, E))], -1).astype(np.float32)
        seq.copy_(torch.from_numpy(a))          # the graph reads the actions at replay time
        graph.replay()
        for j in range(G):
            eager.step(torch.from_numpy(a[j]).cuda())
            want = orc.step(a[j])
        torch.cuda.synchronize()
        assert torch.equal(graphed.get_state(), eager.get_state()), f"replay {r}"
        assert torch.equal(graphed.obs, eager.obs) and torch.equal(graphed.reward, eager.reward)
        assert torch.equal(graphed.get_env_vars(), eager.get_env_vars())
        assert np.array_equal(graphed.obs.cpu().numpy(), want[0]) and np.array_equal(graphed.get_state().cpu().numpy(), orc.state)


@pytest.mark.parametrize("thr", [0.5, 0.0])
def test_offroad_boxes_on_a_random_triangle_soup(oracle, thr):
    """tde_offroad_boxes (flat list of (corner, candidate) pairs) on an adversarial mesh: overlapping triangles from slivers of
    5 cm to 40 m, boxes on, near, far from and off the grid, some absent, a count that is not a multiple of 32 - every value
    equals the oracle's brute force over all triangles."""
    rng = np.random.default_rng(int(thr * 10) + 3)
    n = 300
    c = rng.uniform(-80, 80, (n, 1, 2)); sc = rng.choice([0.05, 0.5, 3.0, 12.0, 40.0], (n, 1, 1))
    v = c + rng.normal(0, 1, (n, 3, 2)) * sc
    road = np.zeros((n, 8), np.float32); road[:, :6] = v.reshape(n, 6)
    ang = rng.uniform(-np.pi, np.pi, n); road[:, 6], road[:, 7] = np.cos(ang), np.sin(ang)
    road = S.subdivide_long_triangles(road, 200.0)
    m = S.MapData(road_tris=road.astype(np.float32), name="soup")
    eng = _engine(S.ScenarioSet([m], [S.make_scenario(0, [[5, 5], [50, 5]], 0, 0, "p")]), 1, 1, offroad_threshold=thr)
    E, A = 301, 19
    st = np.zeros((E, A, 4), np.float32); at = np.zeros((E, A, 4), np.float32)
    st[..., 0:2] = rng.uniform(-130, 130, (E, A, 2)); st[..., 2] = rng.uniform(-np.pi, np.pi, (E, A))
    st[:9, :, 0:2] = rng.uniform(-400, 400, (9, A, 2))                  # far off the grid
    at[..., 0] = rng.uniform(3, 9, (E, A)); at[..., 1] = rng.uniform(1.5, 2.6, (E, A)); at[..., 2] = 1.0
    at[..., 3] = (rng.uniform(0, 1, (E, A)) < 0.85).astype(np.float32)
    got = eng.offroad_boxes(0, torch.from_numpy(st).cuda(), torch.from_numpy(at).cuda()).cpu().numpy()
    want = oracle.offroad_boxes(m.road_tris, thr, st, at)
    assert (want > 0).mean() > 0.2 and (want == 0).mean() > 0.2
    assert np.array_equal(got, want), f"{int((got != want).sum())} of {got.size} differ"
