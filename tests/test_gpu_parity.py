"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle and the
committed golden fixtures.  North-star tolerances: kinematic state within 1e-5 relative; collision /
offroad / termination flags bit-exact away from an epsilon band around contact; birdview pixels
>= 99.9 % identical.  Because both sides follow the same binary32 arithmetic contract the observed
agreement is exact, which the tests also record."""
import numpy as np
import pytest
import torch

from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200._capi import INFO_COLUMNS as IC, PH_INFRACTIONS, PH_KINEMATICS, TdeError, default_config

from golden_util import CASES, load, make_golden

pytestmark = pytest.mark.gpu

STATE_RTOL = 1e-5      # north_star: kinematic state within 1e-5 relative (fp32)
PIXEL_MIN_IDENTICAL = 0.999
EPS_BAND = 1e-4        # metres


def _engine(ss, E, A, **cfg):
    from torchdriveenv_b200.engine import Engine
    return Engine(ss, E, A, device="cuda:0", **cfg)


def _state_close(got, want):
    scale = np.maximum(1.0, np.abs(want))
    return float((np.abs(got - want) / scale).max())


def _rollout_compare(oracle, ss, E, A, steps, seed, strict=True, **cfg):
    eng = _engine(ss, E, A, **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, **cfg), eng.packed)
    eng.reset(seed=seed); orc.reset(seed=seed)
    assert np.array_equal(eng.get_state().cpu().numpy(), orc.state)
    assert np.array_equal(eng.get_env_vars().cpu().numpy(), orc.env_vars)
    assert np.array_equal(eng.render().cpu().numpy(), orc.render())
    rng = np.random.default_rng(seed)
    for k in range(steps):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, r, te, tr, info = eng.step(torch.from_numpy(a).cuda())
        oobs, orr, ote, otr, oinfo = orc.step(a)
        st = eng.get_state().cpu().numpy()
        assert _state_close(st, orc.state) <= STATE_RTOL, f"step {k}"
        assert np.array_equal(te.cpu().numpy(), ote) and np.array_equal(tr.cpu().numpy(), otr), f"step {k}: flags"
        np.testing.assert_allclose(r.cpu().numpy(), orr, rtol=1e-5, atol=1e-5)
        same = float((obs.cpu().numpy() == oobs).mean())
        assert same >= PIXEL_MIN_IDENTICAL, f"step {k}: {same}"
        if strict:
            assert np.array_equal(st, orc.state), f"step {k}: state not bit-exact"
            assert np.array_equal(eng.get_infractions().cpu().numpy(), orc.infractions), f"step {k}: infractions"
            assert np.array_equal(info.cpu().numpy(), oinfo), f"step {k}: info"
            assert np.array_equal(eng.get_env_vars().cpu().numpy(), orc.env_vars), f"step {k}: env vars"
            assert same == 1.0, f"step {k}: obs {same}"
    np.testing.assert_allclose(eng.episode_stats(), orc.stats, rtol=1e-9)
    return eng, orc


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_golden_fixtures(name):
    builder, E, A, steps, over = CASES[name]
    want = load(name)
    eng = _engine(builder(), E, A, **over)
    eng.reset(seed=1234)
    assert np.array_equal(eng.get_state().cpu().numpy(), want["reset_state"])
    assert np.array_equal(eng.get_env_vars().cpu().numpy(), want["reset_vars"])
    assert (eng.render().cpu().numpy() == want["reset_obs"]).mean() >= PIXEL_MIN_IDENTICAL
    for k in range(steps):
        obs, r, te, tr, info = eng.step(torch.from_numpy(want["actions"][k]).cuda())
        assert _state_close(eng.get_state().cpu().numpy(), want["states"][k]) <= STATE_RTOL
        assert np.array_equal(te.cpu().numpy(), want["terminated"][k]) and np.array_equal(tr.cpu().numpy(), want["truncated"][k])
        np.testing.assert_allclose(r.cpu().numpy(), want["reward"][k], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(info.cpu().numpy(), want["info"][k], rtol=1e-5, atol=1e-5)
    assert (obs.cpu().numpy() == want["final_obs"]).mean() >= PIXEL_MIN_IDENTICAL
    assert np.array_equal(obs.cpu().numpy(), want["final_obs"])            # observed: exact
    assert np.array_equal(eng.get_infractions().cpu().numpy(), want["final_infractions"])
    assert np.array_equal(eng.get_env_vars().cpu().numpy(), want["final_vars"])
    np.testing.assert_allclose(eng.episode_stats(), want["stats"], rtol=1e-9)


def test_c1_three_way_single_env(oracle):
    _rollout_compare(oracle, S.three_way(6), 1, 9, steps=200, seed=0)


def test_c2_roundabout_lockstep(oracle):
    _rollout_compare(oracle, S.roundabout(16), 1024, 16, steps=25, seed=1, auto_reset=1)


def test_c3_traffic_lights_birdview(oracle):
    _rollout_compare(oracle, S.traffic_lights(32), 512, 32, steps=30, seed=2, auto_reset=1)


@pytest.mark.parametrize("E,A,n_agents,cfg", [
    (1, 1, 1, dict()),                                              # ego only (EnvConfig.ego_only)
    (13, 5, 3, dict(auto_reset=1)),                                 # ragged: E not a multiple of the block, absent slots
    (37, 64, 64, dict(auto_reset=1, randomize_ego_attributes=1)),   # maximum agents (two lanes per agent slot)
    (9, 33, 33, dict(auto_reset=1, left_handed_coordinates=0)),     # just over one warp of agents, right-handed render
    (21, 8, 8, dict(auto_reset=1, terminated_at_infraction=0, max_environment_steps=7)),
    (16, 8, 8, dict(auto_reset=0, offroad_threshold=0.0, tl_rear_factor=1.0, fov=50.0)),
])
def test_edge_configurations(oracle, E, A, n_agents, cfg):
    _rollout_compare(oracle, S.traffic_lights(n_agents), E, A, steps=20, seed=E + A, **cfg)


def test_scenario_mix_multiple_maps(oracle):
    eng, orc = _rollout_compare(oracle, S.validation_mix(12), 200, 16, steps=20, seed=3, auto_reset=1)
    assert len(set(orc.env_vars[:, 0].tolist())) == 5   # every scenario is drawn
    # restrict envs to scenario sub-ranges
    lo = np.arange(200) % 5; hi = lo + 1
    eng.set_env_scenario_range(lo, hi); orc.set_env_scenario_range(lo, hi)
    eng.reset(seed=5); orc.reset(seed=5)
    v = eng.get_env_vars().cpu().numpy()
    assert np.array_equal(v, orc.env_vars) and np.array_equal(v[:, 0], lo)


def test_maps_without_stop_lines_markings_or_replay(oracle):
    tris = np.array([[0, 0, 60, 0, 60, 8, 1, 0], [0, 0, 60, 8, 0, 8, 1, 0]], np.float32)
    m = S.MapData(road_tris=tris)
    sc = S.ScenarioData(0, np.array([[5, 4], [25, 4], [45, 4]], np.float32), 0.0,
                        np.array([[5, 4, 0, 0], [30, 4, 0, 3]], np.float32), np.array([[5, 2, 0.9], [5, 2, 2.0]], np.float32))
    _rollout_compare(oracle, S.ScenarioSet([m], [sc]), 24, 2, steps=40, seed=4, auto_reset=1)
    # a map with no road at all: offroad is defined as 0
    m0 = S.MapData(road_tris=np.zeros((0, 8), np.float32))
    _rollout_compare(oracle, S.ScenarioSet([m0], [sc]), 8, 2, steps=10, seed=4)


def _triangle_soup(rng, n, scales, extent, width):
    c = rng.uniform(-extent, extent, (n, 1, 2))
    sc = rng.choice(scales, (n, 1, 1))
    v = c + rng.normal(0, 1, (n, 3, 2)) * sc
    out = np.zeros((n, width), np.float32)
    out[:, :6] = v.reshape(n, 6)
    if width == 8:
        ang = rng.uniform(-np.pi, np.pi, n)
        out[:, 6], out[:, 7] = np.cos(ang), np.sin(ang)
    return out


@pytest.mark.parametrize("fov,lh", [(35.0, 1), (12.0, 0), (140.0, 1)])
def test_render_random_triangle_soup(oracle, fov, lh):
    """Adversarial input for the rasteriser (the fixed-point edge accumulators, the quad merging, the top-left rule):
    overlapping random triangles from slivers of 0.05 m to 150 m, random camera poses, three zoom levels - every
    pixel must equal the oracle's per-pixel edge-function test."""
    rng = np.random.default_rng(int(fov) + lh)
    max_edge = min(200.0, 400.0 * fov / 64.0)      # the upload rejects edges beyond 430 px
    road = S.subdivide_long_triangles(_triangle_soup(rng, 500, [0.05, 0.5, 3.0, 20.0, 60.0], 120.0, 8), max_edge)
    mark = S.subdivide_long_triangles(_triangle_soup(rng, 400, [0.05, 0.3, 2.0, 15.0], 120.0, 6), max_edge)
    stop = np.column_stack([rng.uniform(-100, 100, (8, 2)), rng.uniform(0.3, 3, 8), rng.uniform(1, 6, 8), rng.uniform(-3, 3, 8)]).astype(np.float32)
    lights = rng.integers(0, 3, (17, 8)).astype(np.uint8)
    m = S.MapData(road_tris=road.astype(np.float32), mark_tris=mark, stoplines=stop, light_states=lights)
    A = 12
    init = np.column_stack([rng.uniform(-100, 100, (A, 2)), rng.uniform(-3, 3, A), rng.uniform(0, 8, A)]).astype(np.float32)
    attr = np.column_stack([rng.uniform(3, 9, A), rng.uniform(1.5, 2.6, A), rng.uniform(0.8, 2.0, A)]).astype(np.float32)
    sc = S.ScenarioData(0, rng.uniform(-80, 80, (6, 2)).astype(np.float32), 0.3, init, attr)
    E = 384
    cfg = dict(fov=fov, left_handed_coordinates=lh, auto_reset=0)
    eng = _engine(S.ScenarioSet([m], [sc]), E, A, **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, **cfg), eng.packed)
    eng.reset(seed=1); orc.reset(seed=1)
    st = orc.state.copy()
    st[:, 0, 0:2] = rng.uniform(-125, 125, (E, 2)); st[:, 0, 2] = rng.uniform(-np.pi, np.pi, E)
    st[:, 1:, 0:2] += rng.normal(0, 15, (E, A - 1, 2)); st[:, 1:, 2] = rng.uniform(-np.pi, np.pi, (E, A - 1))
    st[: E // 8, 0, 2] = np.round(st[: E // 8, 0, 2] / (np.pi / 2)) * (np.pi / 2)   # axis-aligned cameras: horizontal / vertical edges
    orc.state[...] = st
    eng.set_state(torch.from_numpy(st).cuda())
    got, want = eng.render().cpu().numpy(), orc.render()
    bad = (got != want).reshape(E, -1).any(1)
    assert not bad.any(), f"{int(bad.sum())} envs differ, first {int(np.argmax(bad))}: {(got != want).mean():.2e} of the bytes"
    assert (want.reshape(E, -1).max(1) > 0).mean() > 0.9   # the frames are not empty
    # the same soup through the physics: overlapping triangles of every size in the nearest-candidate grid
    for k in range(4):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, r, te, tr, info = eng.step(torch.from_numpy(a).cuda())
        oobs, orr, ote, otr, oinfo = orc.step(a)
        assert np.array_equal(eng.get_state().cpu().numpy(), orc.state), f"step {k}: state"
        assert np.array_equal(eng.get_infractions().cpu().numpy(), orc.infractions), f"step {k}: infractions"
        assert np.array_equal(info.cpu().numpy(), oinfo) and np.array_equal(obs.cpu().numpy(), oobs), f"step {k}"


def test_granular_entry_points(oracle):
    E, A = 96, 24
    ss = S.traffic_lights(A)
    eng = _engine(ss, E, A)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A), eng.packed)
    eng.reset(seed=6); orc.reset(seed=6)
    rng = np.random.default_rng(6)
    for _ in range(5):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        eng.kinematics(torch.from_numpy(a).cuda()); orc.kinematics(a)      # tde_kinematics: state only
        assert np.array_equal(eng.get_state().cpu().numpy(), orc.state)
    # set_state / compute_infractions / render on an arbitrary state
    st = orc.state.copy()
    st[:, 0, :2] += rng.uniform(-3, 3, (E, 2)).astype(np.float32)
    st[:, 0, 2] += rng.uniform(-1, 1, E).astype(np.float32)
    eng.set_state(torch.from_numpy(st)); orc.state[...] = st
    assert np.array_equal(eng.compute_infractions().cpu().numpy(), orc.compute_infractions())
    assert np.array_equal(eng.render().cpu().numpy(), orc.render())
    at = orc.attr.copy(); at[:, 3:, 3] = 0                                   # hide most NPCs
    eng.set_attributes(torch.from_numpy(at)); orc.attr[...] = at
    assert np.array_equal(eng.compute_infractions().cpu().numpy(), orc.compute_infractions())
    assert np.array_equal(eng.render().cpu().numpy(), orc.render())
    v = orc.env_vars.copy(); v[:, 2] = 3; v[:, 1] = 17
    eng.set_env_vars(torch.from_numpy(v)); orc.env_vars[...] = v
    assert np.array_equal(eng.render().cpu().numpy(), orc.render())          # other target waypoint, other light time
    # phases: kinematics + infractions only (the SimulatorInterface.step path)
    a = np.zeros((E, 2), np.float32)
    eng.step(torch.from_numpy(a).cuda(), render=False, phases=PH_KINEMATICS | PH_INFRACTIONS)
    orc.step(a, render=False, phases=PH_KINEMATICS | PH_INFRACTIONS)
    assert np.array_equal(eng.get_infractions().cpu().numpy(), orc.infractions)
    assert np.array_equal(eng.get_env_vars().cpu().numpy(), orc.env_vars)


def test_collision_flags_bit_exact_outside_epsilon_band(oracle):
    """All-pairs SAT on dense random boxes (config C4's generator at a test size)."""
    for A, size in ((64, 60.0), (32, 30.0), (7, 12.0)):
        st, at = S.scatter_boxes(2048, A, size=size, seed=A, present_p=0.9)
        eng = _engine(S.three_way(0), 1, 3)
        got = eng.collision_boxes(torch.from_numpy(st), torch.from_numpy(at)).cpu().numpy()
        want = oracle.collision_boxes(st, at)
        margins = oracle.collision_margins(st, at)
        outside = margins >= EPS_BAND
        assert outside.mean() > 0.999
        assert np.array_equal(got[outside] > 0, want[outside] > 0)
        assert np.array_equal(got, want)                                     # observed: exact counts everywhere
        assert 0.02 < (want > 0).mean() < 0.9


def test_offroad_exact_vs_bruteforce_oracle(oracle):
    patch = S.scatter_patch(80.0, 10.0)
    ss = S.ScenarioSet([patch, S.build_polyline_map(S.VALIDATION_POLYLINES["roundabout"], "r", ring=dict(center=(0, 0), r_in=13, r_out=26))],
                       [S.make_scenario(0, [[5, 5], [50, 5]], 0, 0, "p")])
    eng = _engine(ss, 2, 1)
    st, at = S.scatter_boxes(4096, 16, size=80.0, seed=8)
    for shift in (0.0, -30.0, 400.0):       # on the grid, straddling its border, far off the grid
        s2 = st.copy(); s2[..., 0] += shift
        got = eng.offroad_boxes(0, torch.from_numpy(s2), torch.from_numpy(at)).cpu().numpy()
        want = oracle.offroad_boxes(patch.road_tris, 0.5, s2, at)
        assert np.array_equal(got, want), shift
    s3 = st.copy(); s3[..., :2] -= 40.0
    got = eng.offroad_boxes(1, torch.from_numpy(s3), torch.from_numpy(at)).cpu().numpy()
    want = oracle.offroad_boxes(ss.maps[1].road_tris, 0.5, s3, at)
    assert np.array_equal(got, want) and 0.1 < (want > 0).mean() < 0.95


def test_full_size_c3_properties(oracle):
    """BASELINE config C3 (16,384 envs x 32 agents with birdview): determinism, shard invariance,
    oracle agreement on env sub-ranges, and invariants that hold for every env."""
    E, A, steps = 16384, 32, 12
    ss = S.traffic_lights(A)
    rng = np.random.default_rng(10)
    acts = np.stack([rng.uniform(-1, 1, (steps, E)), rng.uniform(-0.3, 0.3, (steps, E))], -1).astype(np.float32)

    def run(E_local, offset):
        eng = _engine(ss, E_local, A, auto_reset=1, env_index_offset=offset)
        eng.reset(seed=10)
        out = None
        for k in range(steps):
            o, r, te, tr, info = eng.step(torch.from_numpy(acts[k, offset:offset + E_local]).cuda())
            out = (o.clone(), r.clone(), te.clone(), tr.clone(), info.clone())
        return eng, out

    eng, (obs, rew, term, trunc, info) = run(E, 0)
    eng2, (obs2, rew2, term2, trunc2, info2) = run(E, 0)
    assert torch.equal(obs, obs2) and torch.equal(rew, rew2) and torch.equal(info, info2)          # deterministic
    assert torch.equal(eng.get_state(), eng2.get_state())
    # two shards of 8,192 envs reproduce the single 16,384-env run (what the multi-GPU bench relies on)
    lo_eng, lo_out = run(E // 2, 0)
    hi_eng, hi_out = run(E // 2, E // 2)
    assert torch.equal(torch.cat([lo_out[0], hi_out[0]]), obs) and torch.equal(torch.cat([lo_out[4], hi_out[4]]), info)
    np.testing.assert_allclose(lo_eng.episode_stats() + hi_eng.episode_stats(), eng.episode_stats(), rtol=1e-9)
    # the oracle on three env sub-ranges
    for lo in (0, 7777, E - 64):
        orc = oracle.OracleEnvSet(default_config(num_envs=64, max_agents=A, auto_reset=1, env_index_offset=lo), eng.packed)
        orc.reset(seed=10)
        for k in range(steps):
            oo, orr, ote, otr, oinfo = orc.step(acts[k, lo:lo + 64])
        assert np.array_equal(obs[lo:lo + 64].cpu().numpy(), oo) and np.array_equal(info[lo:lo + 64].cpu().numpy(), oinfo)
        assert np.array_equal(eng.get_state()[lo:lo + 64].cpu().numpy(), orc.state)
    # invariants
    pal = torch.tensor([[0, 0, 0], [128, 128, 128], [255, 255, 255], [0, 200, 0], [230, 200, 0], [220, 0, 0], [0, 170, 255],
                        [60, 90, 220], [250, 120, 0], [200, 220, 255], [255, 230, 150]], dtype=torch.uint8, device="cuda")
    px = obs.permute(0, 2, 3, 1).reshape(-1, 3)
    code = px[:, 0].int() * 65536 + px[:, 1].int() * 256 + px[:, 2].int()
    pcode = pal[:, 0].int() * 65536 + pal[:, 1].int() * 256 + pal[:, 2].int()
    assert bool(torch.isin(code, pcode).all())                                   # only palette colours
    centre = obs[:, :, 32, 32]
    assert bool(((centre == pal[8]).all(1) | (centre == pal[10]).all(1)).all())  # the ego covers the image centre
    inf = info.cpu().numpy()
    any_infr = (inf[:, IC["offroad"]] > 0) | (inf[:, IC["collision"]] > 0) | (inf[:, IC["traffic_light_violation"]] > 0)
    assert np.array_equal(term.cpu().numpy().astype(bool), any_infr)             # is_terminated :413-417
    assert np.array_equal(inf[:, IC["did_reset"]] != 0, (term | trunc).cpu().numpy().astype(bool))
    st = eng.get_state().cpu().numpy()
    stepped = inf[:, IC["did_reset"]] == 0     # a freshly reset ego keeps its unwrapped start heading (set_start_pos :357-361)
    psi = st[stepped][..., 2]
    assert np.isfinite(st).all() and (psi >= -np.pi - 1e-6).all() and (psi < np.pi + 1e-6).all()
    stats = eng.episode_stats()
    assert stats[8] == steps * E and stats[0] == stats[3:6].sum() - 0 or stats[0] <= stats[3:7].sum()


def test_full_size_c4_collision_offroad(oracle):
    """BASELINE config C4: 65,536 envs x 64 agents.  Collision against the oracle in full; offroad on
    the first 2,048 envs (the oracle's brute force is O(triangles)) plus whole-batch properties."""
    E, A = 65536, 64
    st, at = S.scatter_boxes(E, A, size=100.0, seed=12)
    patch = S.scatter_patch(100.0, 10.0)
    eng = _engine(S.ScenarioSet([patch], [S.make_scenario(0, [[5, 5], [50, 5]], 0, 0, "p")]), 1, 1)
    st_d, at_d = torch.from_numpy(st).cuda(), torch.from_numpy(at).cuda()
    col = eng.collision_boxes(st_d, at_d)
    assert np.array_equal(col.cpu().numpy(), oracle.collision_boxes(st, at))
    assert float(col.sum()) % 2 == 0                                             # every hit is seen by both agents
    off = eng.offroad_boxes(0, st_d, at_d)
    assert np.array_equal(off[:2048].cpu().numpy(), oracle.offroad_boxes(patch.road_tris, 0.5, st[:2048], at[:2048]))
    assert bool((off >= 0).all()) and bool(torch.isfinite(off).all())
    # rigid translation by a whole number of tiles leaves the checkerboard, hence offroad, unchanged
    st2 = st.copy(); st2[..., 0] += 20.0
    inside = (st[..., 0] < 70) & (st[..., 0] > 10) & (st[..., 1] > 10) & (st[..., 1] < 90)
    off2 = eng.offroad_boxes(0, torch.from_numpy(st2).cuda(), at_d)
    d = (off2 - off).abs().cpu().numpy()
    assert d[inside].max() < 1e-3


def test_error_behaviour():
    from torchdriveenv_b200.engine import Engine
    ss = S.three_way(0)
    eng = Engine(ss, 4, 3, device="cuda:0")
    with pytest.raises(TdeError, match="TDE_E_STATE"):
        eng.step(torch.zeros(4, 2))                     # step before reset
    eng.reset(seed=0)
    with pytest.raises(TdeError, match="TDE_E_INVAL"):
        eng.set_env_scenario_range(np.zeros(4), np.full(4, 9))
    long_tri = S.MapData(road_tris=np.array([[0, 0, 900, 0, 0, 5, 1, 0]], np.float32))
    with pytest.raises(TdeError, match="TDE_E_SHAPE"):
        Engine(S.ScenarioSet([long_tri], ss.scenarios), 1, 3, device="cuda:0")
    with pytest.raises(TdeError, match="TDE_E_SHAPE"):
        Engine(ss, 1, 2, device="cuda:0") if False else Engine(S.ScenarioSet(ss.maps, ss.scenarios), 1, 65, device="cuda:0")
    eng.close()


def test_host_buffer_entry_point(oracle):
    E, A = 128, 16
    ss = S.roundabout(A)
    eng = _engine(ss, E, A, auto_reset=1)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1), eng.packed)
    eng.reset(seed=14); orc.reset(seed=14)
    rng = np.random.default_rng(14)
    for _ in range(6):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, r, te, tr, info = eng.step_host(a)
        oobs, orr, ote, otr, oinfo = orc.step(a)
        assert np.array_equal(obs, oobs) and np.array_equal(r, orr) and np.array_equal(te, ote) and np.array_equal(info, oinfo)


@pytest.mark.parametrize("E", [4100, 8200])    # 8 and 16 chunks; neither is a multiple of its chunk count
def test_host_buffer_entry_point_chunked_pipeline(E):
    """From 4,096 envs on tde_step_host steps the envs in chunks and sends each chunk's frames back while the
    next chunk is computed: same results as the one-launch device-resident step, statistics included."""
    A = 8
    ss = S.traffic_lights(A)
    eng = _engine(ss, E, A, auto_reset=1)
    ref = _engine(ss, E, A, auto_reset=1)
    eng.reset(seed=21); ref.reset(seed=21)
    rng = np.random.default_rng(21)
    for _ in range(5):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, r, te, tr, info = eng.step_host(a)
        robs, rr, rte, rtr, rinfo = ref.step(torch.from_numpy(a).cuda())
        assert np.array_equal(obs, robs.cpu().numpy()) and np.array_equal(r, rr.cpu().numpy())
        assert np.array_equal(te, rte.cpu().numpy()) and np.array_equal(tr, rtr.cpu().numpy()) and np.array_equal(info, rinfo.cpu().numpy())
        assert torch.equal(eng.get_state(), ref.get_state()) and torch.equal(eng.get_env_vars(), ref.get_env_vars())
    np.testing.assert_allclose(eng.episode_stats(), ref.episode_stats(), rtol=1e-12)


def test_cuda_graph_capture_of_the_step():
    """tde_step only enqueues on the caller's stream, so a step can be captured and replayed."""
    E, A = 256, 16
    eng = _engine(S.roundabout(A), E, A, auto_reset=1)
    ref = _engine(S.roundabout(A), E, A, auto_reset=1)
    eng.reset(seed=15); ref.reset(seed=15)
    act = torch.zeros(E, 2, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        eng.step(act)                                  # warm-up outside capture
        ref.step(act)
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            eng.step(act)
        for k in range(5):
            act.copy_(torch.full((E, 2), 0.1 * k, device="cuda") * torch.tensor([1.0, 0.3], device="cuda"))
            g.replay()
            ref.step(act)
        torch.cuda.current_stream().synchronize()
    torch.cuda.synchronize()
    assert torch.equal(eng.obs, ref.obs) and torch.equal(eng.get_state(), ref.get_state())
