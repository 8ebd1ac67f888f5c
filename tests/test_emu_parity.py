"""Kernel LOGIC checks without a GPU: the product's CUDA sources, compiled for the host-side lockstep emulator
(tests/emu/cuda_emu.h -> tests/emu/libtde_emu.so), against the CPU oracle, bit for bit.

Test infrastructure only.  The emulator runs the same kernel source (warp shuffles, ballots, shared-memory atomics,
named barriers; bulk-async copies as immediate copies) with the threads of a block as fibers on one OS thread; it says
nothing about speed, data races or PTX semantics - the `-m gpu` tests are the parity tests proper.  What it buys: a
logic error in a kernel shows up here, in the CPU suite, before a GPU lease is spent on it."""
import numpy as np
import pytest

from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200._capi import default_config

from emu_engine import EmuEngine


def rollout_compare(oracle, ss, E, A, steps, seed, **cfg):
    eng = EmuEngine(ss, E, A, **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, **cfg), eng.packed)
    eng.reset(seed=seed); orc.reset(seed=seed)
    assert np.array_equal(eng.get_state(), orc.state)
    assert np.array_equal(eng.get_env_vars(), orc.env_vars)
    assert np.array_equal(eng.render(), orc.render()), "reset observation"
    rng = np.random.default_rng(seed)
    for k in range(steps):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, r, te, tr, info = eng.step(a)
        oobs, orr, ote, otr, oinfo = orc.step(a)
        assert np.array_equal(eng.get_state(), orc.state), f"step {k}: state"
        assert np.array_equal(eng.get_infractions(), orc.infractions), f"step {k}: infractions"
        assert np.array_equal(info, oinfo), f"step {k}: info"
        assert np.array_equal(eng.get_env_vars(), orc.env_vars), f"step {k}: env vars"
        assert np.array_equal(te, ote) and np.array_equal(tr, otr), f"step {k}: flags"
        assert np.array_equal(r, orr), f"step {k}: reward"
        assert np.array_equal(obs, oobs), f"step {k}: observation ({float((obs == oobs).mean())} identical)"
    np.testing.assert_allclose(eng.episode_stats(), orc.stats, rtol=1e-9)
    return eng, orc


def test_emu_c3_traffic_lights(oracle):
    rollout_compare(oracle, S.traffic_lights(32), 96, 32, steps=25, seed=2, auto_reset=1)


def test_emu_c2_roundabout(oracle):
    rollout_compare(oracle, S.roundabout(16), 64, 16, steps=20, seed=1, auto_reset=1)


def test_emu_c1_three_way(oracle):
    rollout_compare(oracle, S.three_way(6), 1, 9, steps=120, seed=0)


@pytest.mark.parametrize("E,A,n_agents,cfg", [
    (1, 1, 1, dict()),
    (13, 5, 3, dict(auto_reset=1)),
    (11, 64, 64, dict(auto_reset=1, randomize_ego_attributes=1)),
    (9, 33, 33, dict(auto_reset=1, left_handed_coordinates=0)),
    (21, 8, 8, dict(auto_reset=1, terminated_at_infraction=0, max_environment_steps=7)),
    (16, 8, 8, dict(auto_reset=0, offroad_threshold=0.0, tl_rear_factor=1.0, fov=50.0)),
])
def test_emu_edge_configurations(oracle, E, A, n_agents, cfg):
    rollout_compare(oracle, S.traffic_lights(n_agents), E, A, steps=15, seed=E + A, **cfg)


def test_emu_scenario_mix(oracle):
    rollout_compare(oracle, S.validation_mix(12), 60, 16, steps=15, seed=3, auto_reset=1)


def test_emu_stacked_and_terminal(oracle):
    E, A, n = 24, 8, 3
    ss = S.traffic_lights(A)
    eng = EmuEngine(ss, E, A, auto_reset=1)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1), eng.packed)
    eng.reset(seed=5); orc.reset(seed=5)
    stack, term_obs = eng.new_stack(n), eng.new_stack(n)
    eng.render_stacked(stack, n)
    want = np.zeros_like(stack)
    want[:, 6:] = orc.render()
    assert np.array_equal(stack, want)
    tbuf = np.zeros((E, 3, 64, 64), np.uint8)
    orc.set_terminal_buffer(tbuf)
    rng = np.random.default_rng(5)
    finished = 0
    for k in range(30):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        eng.step_terminal(a, stack, term_obs, n)
        oobs, orr, ote, otr, oinfo = orc.step(a)
        done = (ote | otr).astype(bool)
        prev = want.copy()
        want[:, :6] = prev[:, 3:]
        want[:, 6:] = oobs
        for e in np.nonzero(done)[0]:
            tstack = np.concatenate([prev[e, 3:], tbuf[e]], 0)
            assert np.array_equal(term_obs[e], tstack), f"step {k} env {e}: terminal stack"
            want[e, :6] = 0
            finished += 1
        assert np.array_equal(stack, want), f"step {k}: stack"
        assert np.array_equal(eng.terminated, ote) and np.array_equal(eng.truncated, otr)
    assert finished > 0


def test_emu_collision_and_offroad_boxes(oracle):
    st, at = S.scatter_boxes(24, 64, size=60.0, seed=3)
    patch = S.scatter_patch(60.0, 10.0)
    eng = EmuEngine(S.ScenarioSet([patch], [S.make_scenario(0, [[5, 5], [50, 5]], 0, 0, "p")]), 1, 1)
    assert np.array_equal(eng.collision_boxes(st, at), oracle.collision_boxes(st, at))
    assert np.array_equal(eng.offroad_boxes(0, st, at), oracle.offroad_boxes(patch.road_tris, 0.5, st, at))
    # boxes off the grid (every triangle is a candidate), absent agents, a ragged count (35 boxes: 3 lanes of the second round)
    st3, at3 = S.scatter_boxes(5, 7, size=140.0, seed=8, present_p=0.7)
    st3[..., :2] -= 40.0
    assert np.array_equal(eng.offroad_boxes(0, st3, at3), oracle.offroad_boxes(patch.road_tris, 0.5, st3, at3))
    # a pile-up: more candidate pairs than the compacted pair list holds
    st2, at2 = S.scatter_boxes(6, 64, size=15.0, seed=4)
    got, want = eng.collision_boxes(st2, at2), oracle.collision_boxes(st2, at2)
    assert want.max() > 8
    assert np.array_equal(got, want)


def test_emu_recording_view(oracle):
    E, A = 3, 8
    eng = EmuEngine(S.traffic_lights(A), E, A)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A), eng.packed)
    eng.reset(seed=9); orc.reset(seed=9)
    x, y = [float(v) for v in orc.state[1, 0, :2]]
    got = eng.render_view(1, x, y, 0.3, 120.0, 96, 80)
    assert np.array_equal(got, orc.render_view(1, x, y, 0.3, 120.0, 96, 80))


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
def test_emu_render_random_triangle_soup(oracle, fov, lh):
    """The GPU suite's adversarial rasteriser input (slivers of 0.05 m to triangles of 150 m, overlapping, random and
    axis-aligned cameras, three zoom levels) at a smaller batch: every pixel must equal the oracle's per-pixel test."""
    rng = np.random.default_rng(int(fov) + lh)
    max_edge = min(200.0, 400.0 * fov / 64.0)
    road = S.subdivide_long_triangles(_triangle_soup(rng, 500, [0.05, 0.5, 3.0, 20.0, 60.0], 120.0, 8), max_edge)
    mark = S.subdivide_long_triangles(_triangle_soup(rng, 400, [0.05, 0.3, 2.0, 15.0], 120.0, 6), max_edge)
    stop = np.column_stack([rng.uniform(-100, 100, (8, 2)), rng.uniform(0.3, 3, 8), rng.uniform(1, 6, 8), rng.uniform(-3, 3, 8)]).astype(np.float32)
    lights = rng.integers(0, 3, (17, 8)).astype(np.uint8)
    m = S.MapData(road_tris=road.astype(np.float32), mark_tris=mark, stoplines=stop, light_states=lights)
    A = 12
    init = np.column_stack([rng.uniform(-100, 100, (A, 2)), rng.uniform(-3, 3, A), rng.uniform(0, 8, A)]).astype(np.float32)
    attr = np.column_stack([rng.uniform(3, 9, A), rng.uniform(1.5, 2.6, A), rng.uniform(0.8, 2.0, A)]).astype(np.float32)
    sc = S.ScenarioData(0, rng.uniform(-80, 80, (6, 2)).astype(np.float32), 0.3, init, attr)
    E = 96
    cfg = dict(fov=fov, left_handed_coordinates=lh, auto_reset=0)
    eng = EmuEngine(S.ScenarioSet([m], [sc]), E, A, **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, **cfg), eng.packed)
    eng.reset(seed=1); orc.reset(seed=1)
    st = orc.state.copy()
    st[:, 0, 0:2] = rng.uniform(-125, 125, (E, 2)); st[:, 0, 2] = rng.uniform(-np.pi, np.pi, E)
    st[:, 1:, 0:2] += rng.normal(0, 15, (E, A - 1, 2)); st[:, 1:, 2] = rng.uniform(-np.pi, np.pi, (E, A - 1))
    st[: E // 8, 0, 2] = np.round(st[: E // 8, 0, 2] / (np.pi / 2)) * (np.pi / 2)
    orc.state[...] = st
    eng.set_state(st)
    got, want = eng.render(), orc.render()
    bad = (got != want).reshape(E, -1).any(1)
    assert not bad.any(), f"{int(bad.sum())} envs differ, first {int(np.argmax(bad))}: {(got != want).mean():.2e} of the bytes"
    assert (want.reshape(E, -1).max(1) > 0).mean() > 0.9
    for k in range(2):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, r, te, tr, info = eng.step(a)
        oobs, orr, ote, otr, oinfo = orc.step(a)
        assert np.array_equal(eng.get_state(), orc.state), f"step {k}: state"
        assert np.array_equal(eng.get_infractions(), orc.infractions), f"step {k}: infractions"
        assert np.array_equal(info, oinfo) and np.array_equal(obs, oobs), f"step {k}"


@pytest.mark.parametrize("E,A,ss", [(96, 32, "traffic_lights"), (40, 64, "traffic_lights"), (64, 16, "roundabout"), (30, 16, "mix")])
def test_emu_physics_with_staged_map_tables(oracle, E, A, ss):
    """cfg.stage_map_tables = 1: the physics kernel copies the map tables into shared memory with bulk-async copies
    (emulated as copies gated by the mbarrier) and answers SAFE corners / followed lanes from the per-cell summary; same
    results as the launch that reads the tables from global memory.  The five-map mix does not fit and falls back."""
    sets = dict(traffic_lights=lambda: S.traffic_lights(A), roundabout=lambda: S.roundabout(A), mix=lambda: S.validation_mix(12))
    rollout_compare(oracle, sets[ss](), E, A, steps=12, seed=11, auto_reset=1, stage_map_tables=1)


@pytest.mark.parametrize("policy", ["1", "2"])
def test_emu_is_independent_of_warp_scheduling(policy):
    """The emulator's adversarial schedules (one end of the block always runs first, so a warp gets as far ahead of the
    others as the barriers allow): results must not depend on which warp of a group or CTA runs first."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, TDE_EMU_SCHED=policy)
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", os.path.join(here, "test_emu_parity.py"), "-k",
                        "c3_traffic or stacked or scenario_mix or edge_configurations or staged"], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("NS,T", [(3, 7), (4, 3), (2, 5)])
def test_emu_rollout_scatter_fills_the_buffer_like_vec_frame_stack(oracle, NS, T):
    """tde_step_rollout_scatter: every slot of a [T + 1, E, 3 n, 64, 64] rollout buffer must hold what VecFrameStack
    would have produced (oracle frames stacked oldest first, zeros before a restart), over three rollouts with the
    carry-over of the last slot - without the kernel ever reading a frame."""
    from emu_engine import _aligned
    E, A = 20, 6
    ss = S.traffic_lights(A)
    cfg = dict(auto_reset=1, max_environment_steps=9)
    eng = EmuEngine(ss, E, A, **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, **cfg), eng.packed)
    eng.reset(seed=13); orc.reset(seed=13)
    buf = _aligned((T + 1, E, 3 * NS, 64, 64), np.uint8)
    eng.render_stacked(buf[0], NS)
    frames, age = [orc.render()], np.zeros(E, np.int64)

    def want_stack():
        w = np.zeros((E, 3 * NS, 64, 64), np.uint8)
        for slot in range(NS):
            back = NS - 1 - slot
            if back < len(frames):
                have = age >= back
                w[have, 3 * slot:3 * slot + 3] = frames[-1 - back][have]
        return w

    def seed_older():
        for j in range(1, min(NS, T + 1)):
            buf[j][:, : 3 * (NS - j)] = buf[0][:, 3 * j:]

    rng = np.random.default_rng(13)
    n_done = 0
    for r in range(3):
        if r:
            buf[0] = buf[T]
        assert np.array_equal(buf[0], want_stack())
        seed_older()
        for t in range(T):
            a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
            eng.step_rollout_scatter(a, buf, t, NS)
            oobs, orr, ote, otr, oinfo = orc.step(a)
            d = (ote | otr).astype(bool)
            frames.append(oobs)
            age = np.where(d, 0, age + 1)
            assert np.array_equal(buf[t + 1], want_stack()), f"rollout {r} step {t}"
            assert np.array_equal(eng.reward, orr) and np.array_equal(eng.info, oinfo)
            n_done += int(d.sum())
    assert n_done > 5


@pytest.mark.parametrize("E,steps", [(37, 12), (300, 4)])
def test_emu_class_image_and_compact_host_step(oracle, monkeypatch, E, steps):
    """tde_render_classes is the oracle's class image, nibble-packed; tde_step_host's default path (the class
    image crosses PCIe, host threads apply the palette) fills the caller's buffer with the same bytes as cfg.host_obs_rgb = 1,
    with one chunk and with several, with the default and with a caller-set palette; a handful of envs is expanded by the
    calling thread, 300 by the pool of host threads."""
    A = 16
    ss = S.validation_mix(12)
    cfg = dict(auto_reset=1)
    rgb = EmuEngine(ss, E, A, host_obs_rgb=1, **cfg)
    cmp_ = EmuEngine(ss, E, A, **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, **cfg), rgb.packed)
    pal = np.random.default_rng(5).integers(0, 256, (11, 3)).astype(np.uint8)
    rng = np.random.default_rng(9)
    for eng in (rgb, cmp_):
        eng.reset(seed=4)
    orc.reset(seed=4)
    assert np.array_equal(cmp_.render_classes(), orc.render_classes())
    for k in range(steps):
        if k == steps // 2:
            for eng in (rgb, cmp_, orc):
                eng.set_palette(pal)
        monkeypatch.setenv("TDE_HOST_CHUNKS", "1" if k % 2 else "5")
        monkeypatch.setenv("TDE_HOST_SIMD", ("scalar", "avx2", "avx512")[k % 3])   # every expansion loop the CPU has
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        want = [np.array(x, copy=True) for x in rgb.step_host(a)]
        got = cmp_.step_host(a)
        ref = orc.step(a)
        for name, w, g, o in zip(("obs", "reward", "terminated", "truncated", "info"), want, got, ref):
            assert np.array_equal(w, g), f"step {k}: {name} differs between the RGB and the compact host step"
            assert np.array_equal(g, o), f"step {k}: {name} differs from the oracle"
        assert np.array_equal(cmp_.render_classes(), orc.render_classes()), f"step {k}: class image"


@pytest.mark.parametrize("NS,R", [(3, 4), (2, 2), (4, 7)])
def test_emu_stacked_ring_is_vec_frame_stack(oracle, NS, R):
    """tde_step_stacked_ring: slot pos of the ring is VecFrameStack's observation of the step (the oracle's frames stacked,
    zeros before an episode's first frame), and the slots of the last R - NS steps are still intact."""
    from emu_engine import _aligned
    E, A = 21, 6
    ss = S.traffic_lights(A)
    cfg = dict(auto_reset=1, max_environment_steps=7)
    eng = EmuEngine(ss, E, A, **cfg)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, **cfg), eng.packed)
    eng.reset(seed=5); orc.reset(seed=5)
    ring = _aligned((R, E, 3 * NS, 64, 64), np.uint8)
    eng.render_stacked(ring[0], NS)
    for j in range(1, min(NS, R)):
        ring[j][:, : 3 * (NS - j)] = ring[0][:, 3 * j:]
    stack = np.zeros((E, 3 * NS, 64, 64), np.uint8)
    stack[:, -3:] = orc.render()
    assert np.array_equal(ring[0], stack)
    history = [stack.copy()]
    rng = np.random.default_rng(1)
    n_done = 0
    for k in range(1, 40):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, r, te, tr, info = eng.step_stacked_ring(a, ring, k, NS)
        oobs, orr, ote, otr, oinfo = orc.step(a)
        done = (ote | otr).astype(bool)
        n_done += int(done.sum())
        stack = np.concatenate((stack[:, 3:], oobs), 1)
        stack[done, : 3 * (NS - 1)] = 0
        history.append(stack.copy())
        assert np.array_equal(obs, stack), f"step {k}"
        assert np.array_equal(r, orr) and np.array_equal(info, oinfo)
        for back in range(1, R - NS + 1):       # older observations that must have survived
            if k - back >= 0:
                assert np.array_equal(ring[(k - back) % R], history[k - back]), f"step {k}: the observation of step {k - back} was overwritten"
    assert n_done > 10


@pytest.mark.parametrize("thr", [0.5, 0.0])
def test_emu_offroad_boxes_on_a_random_triangle_soup(oracle, thr):
    """The stateless offroad kernel (flat list of (corner, candidate) pairs) on an adversarial mesh - overlapping triangles
    from slivers of 5 cm to 40 m, so cells list many candidates and several containing triangles - with boxes on, near, far
    from and off the grid, some absent, in a count that is not a multiple of 32: every value equals the oracle's brute
    force over all triangles."""
    rng = np.random.default_rng(int(thr * 10) + 3)
    road = S.subdivide_long_triangles(_triangle_soup(rng, 300, [0.05, 0.5, 3.0, 12.0, 40.0], 80.0, 8), 200.0)
    m = S.MapData(road_tris=road.astype(np.float32), name="soup")
    eng = EmuEngine(S.ScenarioSet([m], [S.make_scenario(0, [[5, 5], [50, 5]], 0, 0, "p")]), 1, 1, offroad_threshold=thr)
    E, A = 7, 19
    st = np.zeros((E, A, 4), np.float32); at = np.zeros((E, A, 4), np.float32)
    st[..., 0:2] = rng.uniform(-130, 130, (E, A, 2)); st[..., 2] = rng.uniform(-np.pi, np.pi, (E, A))
    st[0, :, 0:2] = rng.uniform(-400, 400, (A, 2))                      # far off the grid
    at[..., 0] = rng.uniform(3, 9, (E, A)); at[..., 1] = rng.uniform(1.5, 2.6, (E, A)); at[..., 2] = 1.0
    at[..., 3] = (rng.uniform(0, 1, (E, A)) < 0.85).astype(np.float32)
    got, want = eng.offroad_boxes(0, st, at), oracle.offroad_boxes(m.road_tris, thr, st, at)
    assert (want > 0).mean() > 0.2 and (want == 0).mean() > 0.2
    assert np.array_equal(got, want), f"{int((got != want).sum())} of {got.size} differ"
