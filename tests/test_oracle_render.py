"""Pins the oracle's birdview: exact equality with an independent numpy restatement of the same
fixed-point rule, and looser agreement with float64 (no snapping) and cv2.fillPoly renderings."""
import numpy as np
import pytest

from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200._capi import default_config

from raster_ref import camera, raster_cv2, raster_fixed, raster_float64, world_primitives


def _envs(oracle, ss, E, A, seed, steps, **cfg):
    c = default_config(num_envs=E, max_agents=A, **cfg)
    packed = ss.pack(A)
    env = oracle.OracleEnvSet(c, packed)
    env.reset(seed=seed)
    rng = np.random.default_rng(seed)
    for _ in range(steps):
        env.step(np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1), render=False)
    return env, c, packed


@pytest.mark.parametrize("name,builder,A,lh", [("three_way", S.three_way, 9, 1), ("traffic_lights", lambda: S.traffic_lights(24), 24, 1),
                                                 ("roundabout", lambda: S.roundabout(12), 12, 0)])
def test_oracle_render_vs_independent_rasterisers(oracle, name, builder, A, lh):
    E = 6
    env, cfg, packed = _envs(oracle, builder(), E, A, seed=5, steps=7, auto_reset=1, left_handed_coordinates=lh)
    cls = env.render_classes()
    obs = env.render()
    pal = np.array([[0, 0, 0], [128, 128, 128], [255, 255, 255], [0, 200, 0], [230, 200, 0], [220, 0, 0], [0, 170, 255],
                    [60, 90, 220], [250, 120, 0], [200, 220, 255], [255, 230, 150]], np.uint8)
    same64, samecv = [], []
    for e in range(E):
        prims = world_primitives(packed, cfg, env.state[e], env.attr[e], env.env_vars[e], oracle.sincos)
        cam = camera(cfg, env.state[e, 0], oracle.sincos)
        ref = raster_fixed(prims, cam)
        assert np.array_equal(ref, cls[e]), f"{name} env {e}: {(ref != cls[e]).sum()} pixels differ from the fixed-point restatement"
        assert np.array_equal(obs[e], pal[cls[e]].transpose(2, 0, 1))
        same64.append((raster_float64(prims, cam) == cls[e]).mean())
        samecv.append((raster_cv2(prims, cam) == cls[e]).mean())
        assert (cls[e] == 8).sum() > 10 and cls[e, 32, 32] in (8, 9, 10)   # the ego sits at the image centre
    assert np.mean(same64) > 0.985, same64   # differences only where snapping / tie rules matter
    assert np.mean(samecv) > 0.93, samecv    # fillPoly also paints boundary pixels


def test_render_orientation_and_handedness(oracle):
    """Ego faces +x of the image; a point ahead of the ego lands right of centre; a point to the ego's
    left (+y in a right-handed frame) lands above centre right-handed and below it left-handed."""
    ss = S.three_way(0)
    for lh, row_sign in ((1, +1), (0, -1)):
        cfg = default_config(num_envs=1, max_agents=3, left_handed_coordinates=lh)
        env = oracle.OracleEnvSet(cfg, ss.pack(3))
        env.reset(seed=0)
        env.state[0, 0] = [0, 0, 0.5, 0]
        env.attr[0, 1:, 3] = 0
        env.attr[0, 1, :] = [4, 2, 1, 1]
        c, s = np.cos(0.5), np.sin(0.5)
        env.state[0, 1] = [8 * c, 8 * s, 0.5, 0]           # 8 m straight ahead
        cls = env.render_classes()[0]
        ys, xs = np.nonzero(cls == 7)
        assert xs.mean() > 40 and abs(ys.mean() - 32) < 1.5
        env.state[0, 1] = [-8 * s, 8 * c, 0.5, 0]          # 8 m to the left (+90 deg)
        cls = env.render_classes()[0]
        ys, xs = np.nonzero(cls == 7)
        assert abs(xs.mean() - 32) < 1.5 and (ys.mean() - 32) * row_sign > 8
        env.close()


def test_render_fill_rule_shared_edges(oracle):
    """Two triangles sharing an edge cover every pixel exactly once-or-more (no cracks), and an
    axis-aligned square covers exactly the pixel centres inside it (top-left rule on the boundary)."""
    tris = np.array([[0, 0, 20, 0, 20, 20, 1, 0], [0, 0, 20, 20, 0, 20, 1, 0]], np.float32)
    m = S.MapData(road_tris=tris)
    sc = S.ScenarioData(0, np.array([[10, 10], [15, 10]], np.float32), 0.0, np.array([[10, 10, 0, 0]], np.float32), np.array([[0.2, 0.2, 1]], np.float32))
    cfg = default_config(num_envs=1, max_agents=1, fov=64.0)   # 1 px per metre
    env = oracle.OracleEnvSet(cfg, S.ScenarioSet([m], [sc]).pack(1))
    env.reset(seed=0)
    for psi in (0.0, 0.3, 1.1, 2.5):
        env.state[0, 0] = [10.25, 10.25, psi, 0]
        cls = env.render_classes()[0]
        road = (cls >= 1)
        # the 20 m square is 400 px^2; pixel-centre sampling gives 400 +- boundary effects
        assert abs(int(road.sum()) - 400) <= 12
        # no crack along the shared diagonal: the covered region has no holes
        filled = road.copy()
        assert (filled[1:-1, 1:-1] | ~(filled[:-2, 1:-1] & filled[2:, 1:-1] & filled[1:-1, :-2] & filled[1:-1, 2:])).all()
    env.state[0, 0] = [10.0, 10.0, 0.0, 0]   # square spans pixel x in [22, 42): centres 22.5 .. 41.5
    cls = env.render_classes()[0]
    ys, xs = np.nonzero(cls >= 1)
    assert xs.min() == 22 and xs.max() == 41 and ys.min() == 22 and ys.max() == 41


def test_recording_view_at_observation_settings_is_the_observation(oracle):
    """orc_render_view (twin of tde_render_view, the BirdviewRecordingWrapper frame) with the camera on the ego,
    64 x 64 px over the observation's fov draws from the raw triangles through the wide-range fixed-point path:
    it must give the observation itself, for both handednesses."""
    from torchdriveenv_b200 import scenarios as S
    from torchdriveenv_b200._capi import default_config
    for lh in (1, 0):
        ss = S.validation_mix(8); A = 10
        orc = oracle.OracleEnvSet(default_config(num_envs=10, max_agents=A, left_handed_coordinates=lh), ss.pack(A))
        orc.reset(seed=5)
        rng = np.random.default_rng(5)
        for _ in range(3):
            orc.step(np.stack([rng.uniform(-1, 1, 10), rng.uniform(-0.3, 0.3, 10)], 1).astype(np.float32))
        obs = orc.render()
        for e in range(10):
            x, y, psi = (float(v) for v in orc.state[e, 0, :3])
            assert np.array_equal(orc.render_view(e, x, y, psi, 35.0, 64, 64), obs[e]), (lh, e)


def test_recording_view_whole_map(oracle):
    """A whole-map frame: every class that exists in the scene shows up, the image is mostly background, and a
    2x zoom on the same centre keeps the centre pixel's class."""
    from torchdriveenv_b200 import scenarios as S
    from torchdriveenv_b200._capi import default_config
    ss = S.traffic_lights(12); A = 12
    orc = oracle.OracleEnvSet(default_config(num_envs=1, max_agents=A), ss.pack(A))
    orc.reset(seed=2)
    x, y = (float(v) for v in orc.state[0, 0, :2])
    far = orc.render_view(0, x, y, 0.0, 400.0, 512, 384)
    near = orc.render_view(0, x, y, 0.0, 200.0, 512, 384)
    assert far.shape == (3, 384, 512)
    colours = {tuple(c) for c in far.reshape(3, -1).T}
    assert (128, 128, 128) in colours and (255, 255, 255) in colours and (250, 120, 0) in colours   # road, markings, ego
    assert (far.sum(0) == 0).mean() > 0.8
    assert np.array_equal(far[:, 192, 256], near[:, 192, 256])
    with pytest.raises(ValueError):
        orc.render_view(0, x, y, 0.0, 0.0, 64, 64)


def test_oracle_render_triangle_soup_vs_fixed_point_restatement(oracle):
    """The adversarial input of tests/test_gpu_parity.py::test_render_random_triangle_soup (overlapping random triangles
    from slivers to 60 m, random cameras) through the oracle and through the independent numpy restatement of the
    fixed-point rule: every pixel equal.  Pins the checker the CUDA rasteriser is compared with on that input."""
    rng = np.random.default_rng(3)

    def soup(n, scales, width):
        c = rng.uniform(-60, 60, (n, 1, 2)); sc = rng.choice(scales, (n, 1, 1))
        out = np.zeros((n, width), np.float32)
        out[:, :6] = (c + rng.normal(0, 1, (n, 3, 2)) * sc).reshape(n, 6)
        if width == 8:
            out[:, 6] = 1.0
        return out
    road = S.subdivide_long_triangles(soup(160, [0.05, 0.5, 3.0, 20.0], 8), 200.0)
    mark = S.subdivide_long_triangles(soup(120, [0.05, 0.3, 2.0, 10.0], 6), 200.0)
    m = S.MapData(road_tris=road, mark_tris=mark)
    A, E = 4, 10
    init = np.column_stack([rng.uniform(-50, 50, (A, 2)), rng.uniform(-3, 3, A), rng.uniform(0, 8, A)]).astype(np.float32)
    attr = np.column_stack([rng.uniform(3, 9, A), rng.uniform(1.5, 2.6, A), rng.uniform(0.8, 2.0, A)]).astype(np.float32)
    sc = S.ScenarioData(0, rng.uniform(-40, 40, (4, 2)).astype(np.float32), 0.3, init, attr)
    ss = S.ScenarioSet([m], [sc])
    for lh in (1, 0):
        cfg = default_config(num_envs=E, max_agents=A, left_handed_coordinates=lh)
        packed = ss.pack(A)
        env = oracle.OracleEnvSet(cfg, packed)
        env.reset(seed=2)
        env.state[:, 0, 0:2] = rng.uniform(-65, 65, (E, 2)); env.state[:, 0, 2] = rng.uniform(-np.pi, np.pi, E)
        env.state[:2, 0, 2] = [0.0, np.pi / 2]          # axis-aligned cameras
        cls = env.render_classes()
        for e in range(E):
            prims = world_primitives(packed, cfg, env.state[e], env.attr[e], env.env_vars[e], oracle.sincos)
            ref = raster_fixed(prims, camera(cfg, env.state[e, 0], oracle.sincos))
            assert np.array_equal(ref, cls[e]), f"lh {lh} env {e}: {(ref != cls[e]).sum()} pixels differ"
        assert (cls > 0).mean() > 0.2
