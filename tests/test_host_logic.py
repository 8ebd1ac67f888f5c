"""Host-side logic that needs no GPU: scenario tables, file loaders, config mirror, sharding."""
import json
import os

import numpy as np
import pytest
import yaml

from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200.distributed import shard_range, summarize
from torchdriveenv_b200.roofline import bytes_per_env_step, c4_bytes_per_env


def test_pack_layout_and_padding():
    ss = S.validation_mix(6)
    A = 12
    p = ss.pack(A)
    Nm, Ns = len(ss.maps), len(ss.scenarios)
    assert p["map_tri_offset"].shape == (Nm + 1,) and p["map_tri_offset"][-1] == p["road_tris"].shape[0]
    assert p["road_tris"].dtype == np.float32 and p["road_tris"].shape[1] == 8
    assert p["agent_init"].shape == (Ns, A, 4) and p["agent_attr"].shape == (Ns, A, 3)
    assert p["replay_states"].shape == (p["scen_replay_offset"][-1], A * 4)
    assert p["replay_mask"].shape == (p["scen_replay_offset"][-1], A)
    assert (p["replay_mask"][:, 0] == 0).all()                       # the ego is never replayed
    for k, sc in enumerate(ss.scenarios):
        n = sc.agent_init.shape[0]
        assert p["scen_num_agents"][k] == n
        np.testing.assert_array_equal(p["agent_init"][k, :n], sc.agent_init)
        assert (p["agent_init"][k, n:] == 0).all()
    L = ss.maps[4].stoplines.shape[0]
    assert L > 0 and p["map_light_offset"][5] - p["map_light_offset"][4] == p["map_light_period"][4] * L
    with pytest.raises(ValueError):
        ss.pack(3)


def test_lane_directions_are_unit_and_follow_the_polyline():
    m = S.build_polyline_map(S.VALIDATION_POLYLINES["chicken"], "c")
    d = m.road_tris[:, 6:8]
    np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
    # this polyline runs towards -y: the ego lane points down, the oncoming lane up
    assert (d[:, 1] < -0.9).sum() > 10 and (d[:, 1] > 0.9).sum() > 10


def test_subdivide_long_triangles():
    t = np.array([[0, 0, 1000, 0, 0, 10, 1, 0]], np.float32)
    out = S.subdivide_long_triangles(t, 200.0)
    v = out[:, :6].reshape(-1, 3, 2)
    e = np.linalg.norm(v - np.roll(v, -1, axis=1), axis=2)
    assert e.max() <= 200.0 + 1e-3 and out.shape[0] >= 5
    area = 0.5 * np.abs((v[:, 1, 0] - v[:, 0, 0]) * (v[:, 2, 1] - v[:, 0, 1]) - (v[:, 1, 1] - v[:, 0, 1]) * (v[:, 2, 0] - v[:, 0, 0]))
    assert abs(area.sum() - 5000.0) < 1.0
    assert (out[:, 6] == 1).all()


def test_replay_rollout_follows_the_lane():
    ss = S.traffic_lights(8)
    sc = ss.scenarios[0]
    assert sc.replay_states.shape == (200, 8, 4) and sc.replay_mask[:, 1:].all() and not sc.replay_mask[:, 0].any()
    step = np.linalg.norm(np.diff(sc.replay_states[:, 1:, :2], axis=0), axis=2)
    speed = sc.replay_states[0, 1:, 3]
    moving = step[:100]
    assert np.all(moving <= speed[None] * 0.1 * 1.05 + 1e-3)


def test_shard_range_partitions():
    for total, ws in ((16384, 8), (10, 3), (5, 8), (65536, 4)):
        spans = [shard_range(total, r, ws) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def test_summarize_and_roofline_formula():
    s = np.zeros(16); s[:9] = [10, 50, 400, 4, 2, 1, 3, 25, 400]
    out = summarize(s)
    assert out["mean_return"] == 5 and out["mean_length"] == 40 and out["offroad_rate"] == 0.4 and out["success_rate"] == 0.3
    assert bytes_per_env_step(16, False) == 1010 and bytes_per_env_step(32, True) == 14274
    assert c4_bytes_per_env(64) == 1856


def test_env_config_defaults_match_reference():
    from torchdriveenv_b200.gym_env import EnvConfig, Scenario, WaypointSuite
    c = EnvConfig()
    assert (c.ego_only, c.max_environment_steps, c.frame_stack, c.waypoint_bonus, c.heading_penalty, c.distance_bonus,
            c.distance_cutoff, c.use_background_traffic, c.terminated_at_infraction, c.seed, c.render_mode, c.video_res,
            c.video_fov, c.device) == (False, 200, 3, 100., 25., 1., 0.5, True, True, None, "rgb_array", 1024, 500, None)
    assert c.simulator.left_handed_coordinates and c.simulator.highlight_ego_vehicle
    assert Scenario().agent_states is None and WaypointSuite().locations is None


def test_yaml_and_json_loaders(tmp_path):
    from torchdriveenv_b200 import env_utils
    from torchdriveenv_b200.gym_env import EnvConfig, scenario_set_from_suite
    suite = dict(locations=["Town01", "Town03"],
                 waypoint_suite=[[[0.0, 0.0], [10.0, 0.0], [20.0, 5.0]], [[5.0, 5.0], [5.0, 25.0]]],
                 car_sequence_suite=[{1: [[30.0 + t, 0.0, 0.0, 0.0] for t in range(5)]}, {}],
                 scenarios=[dict(agent_states=[[30.0, 0.0, 0.0, 0.0]], agent_attributes=[[5.0, 2.0, 2.0]], recurrent_states=[[0.0] * 4]), None])
    p = tmp_path / "suite.yml"
    p.write_text(yaml.safe_dump(suite))
    data = env_utils.load_waypoint_suite_data(str(p))
    assert data.locations == ["Town01", "Town03"] and data.scenarios[1] is None and data.scenarios[0].agent_attributes == [[5.0, 2.0, 2.0]]
    ss = scenario_set_from_suite(EnvConfig(), data, n_background=2, seed=0)
    assert len(ss.maps) == 2 and ss.scenarios[0].agent_init.shape == (4, 4)
    assert ss.scenarios[0].replay_mask[:5, 1].all() and not ss.scenarios[0].replay_mask[5:, 1].any()
    np.testing.assert_allclose(ss.scenarios[0].replay_states[3, 1], [33, 0, 0, 0])
    assert ss.scenarios[1].agent_init.shape == (3, 4)
    ego_only = scenario_set_from_suite(EnvConfig(ego_only=True), data, n_background=2)
    assert ego_only.scenarios[0].agent_init.shape == (1, 4)
    cfgp = tmp_path / "env.yml"
    cfgp.write_text(yaml.safe_dump(dict(distance_cutoff=0.25, max_environment_steps=100, simulator=dict(offroad_threshold=0.3))))
    cfg = env_utils.load_env_config(str(cfgp))
    assert cfg.distance_cutoff == 0.25 and cfg.max_environment_steps == 100 and cfg.simulator.offroad_threshold == 0.3
    # scenario-builder JSON (env_utils.py:31-105)
    d = tmp_path / "labeled"; d.mkdir()
    st = lambda x, y: dict(center=dict(x=x, y=y), orientation=0.5)
    js = dict(individual_suggestions={"0": dict(states=[st(0, 0), st(10, 0)])},
              predetermined_agents={"1": dict(states={"0": st(20, 0)}, static_attributes=dict(length=5, width=2, rear_axis_offset=1.5, max_speed=0)),
                                    "2": dict(states={"0": st(30, 0), "1": st(31, 0)}, static_attributes=dict(length=4.5, width=1.9, rear_axis_offset=1.4))})
    (d / "carla_Town02_x.json").write_text(json.dumps(js))
    lab = env_utils.load_labeled_data(str(d))
    assert lab.locations == ["Town02"] and lab.waypoint_suite[0] == [[0, 0], [10, 0]]
    assert len(lab.car_sequence_suite[0][1]) == 200 and len(lab.car_sequence_suite[0][2]) == 2
    assert lab.scenarios[0].agent_attributes[1] == [4.5, 1.9, 1.4]


@pytest.mark.skipif(not os.path.exists("/root/reference/torchdriveenv/data/validation_cases.yml"), reason="reference checkout not present")
def test_reference_validation_suite_loads():
    from torchdriveenv_b200 import env_utils
    from torchdriveenv_b200.gym_env import EnvConfig, scenario_set_from_suite
    data = env_utils.load_waypoint_suite_data("/root/reference/torchdriveenv/data/validation_cases.yml")
    assert data.locations == ["Town07", "Town07", "Town03", "Town03", "Town01"]
    for name, poly in zip(["three_way", "parked_car", "chicken", "roundabout", "traffic_lights"], data.waypoint_suite):
        np.testing.assert_allclose(np.asarray(poly), np.asarray(S.VALIDATION_POLYLINES[name]), atol=1e-3)
    ss = scenario_set_from_suite(EnvConfig(), data)
    assert ss.scenarios[1].replay_states.shape[0] == 300   # parked-car case: 300-step stationary replay
    ss.pack(ss.max_agents())


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from torchdriveenv_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(S.three_way(0), 1)


def test_background_traffic_loader_and_light_program(tmp_path):
    """Background-traffic JSON schema of the reference (gym_env.py:200-233) and the light-controller
    program compiler (gym_env.py:181-189, 290-291)."""
    import json, os
    from torchdriveenv_b200 import env_utils as U, scenarios as S
    from torchdriveenv_b200.gym_env import EnvConfig, WaypointSuite, scenario_set_from_suite
    agents = [dict(center=dict(x=10.0 * k, y=-5.0), orientation=0.1 * k, speed=3.0 + k) for k in range(30)]
    attrs = [dict(length=4.5 + 0.01 * k, width=1.9, rear_axis_offset=1.2) for k in range(30)]
    doc = dict(location="carla:Town03", agent_density=20, random_seed=7, agent_states=agents, agent_attributes=attrs,
               recurrent_states=[dict(packed=[0.0] * 4)] * 30)
    d = tmp_path / "background_traffic"; d.mkdir()
    (d / "carla_Town03_20_7.json").write_text(json.dumps(doc))
    doc2 = dict(doc, agent_density=90)                 # 30 + 90 >= 100: never chosen (:214)
    (d / "carla_Town03_90_8.json").write_text(json.dumps(doc2))
    (d / "carla_Town01_20_9.json").write_text(json.dumps(doc))
    bt = U.pick_background_traffic(str(d), "Town03")
    assert bt["agent_density"] == 20 and bt["states"].shape == (30, 4) and bt["attributes"].shape == (30, 3)
    assert U.pick_background_traffic(str(d), "Town07") is None
    st, at = U.background_agents_for_start(bt, (0.0, -5.0))
    assert len(st) == 19 and (np.hypot(st[:, 0], st[:, 1] + 5.0) > 100).all()     # x = 110 .. 290
    np.testing.assert_allclose(at[0], [4.5 + 0.11, 1.9, 1.2], rtol=1e-6)
    suite = WaypointSuite(locations=["Town03"], waypoint_suite=[[[0, -5], [20, -5], [40, -5]]], car_sequence_suite=[None], scenarios=[None])
    ss = scenario_set_from_suite(EnvConfig(), suite, background_traffic=bt)
    assert ss.scenarios[0].agent_init.shape == (20, 4)                             # ego + 19 kept agents
    np.testing.assert_allclose(ss.scenarios[0].agent_init[1], [110.0, -5.0, 1.1, 14.0], rtol=1e-6)
    ref_dir = "/root/reference/torchdriveenv/resources/background_traffic"        # the real files, where the checkout exists
    if os.path.isdir(ref_dir):
        names = [n for n in os.listdir(ref_dir) if n.endswith(".json")]
        assert names
        for n in names:
            b = U.load_background_traffic(os.path.join(ref_dir, n))
            assert b["states"].shape[1] == 4 and b["states"].shape[0] == b["attributes"].shape[0] and np.isfinite(b["states"]).all()
    # light program: 2 s green, 0.5 s yellow, 1.5 s red on light 0; light 1 opposite
    sched = S.compile_light_program([(2.0, ["green", "red"]), (0.5, {0: "yellow"}), (1.5, {0: "red", 1: "green"})], 2, dt=0.1)
    assert sched.shape == (40, 2) and sched.dtype == np.uint8
    assert (sched[:20, 0] == S.LIGHT_GREEN).all() and (sched[20:25, 0] == S.LIGHT_YELLOW).all() and (sched[25:, 0] == S.LIGHT_RED).all()
    assert (sched[:25, 1] == S.LIGHT_RED).all() and (sched[25:, 1] == S.LIGHT_GREEN).all()
    with pytest.raises(ValueError):
        S.compile_light_program([(1.0, ["green"])], 2)


def test_training_mix_uses_the_reference_training_polylines():
    """Config C5's scenario mix is built on the reference's own 100 training polylines (training_cases.yml, packaged as
    torchdriveenv_b200/data/training_cases.json); beyond 100, synthetic polylines with the same statistics follow
    (5-20 waypoints, spacing 12.9-15.0 m)."""
    import os
    from torchdriveenv_b200 import env_utils as U
    polys = S.reference_training_polylines()
    assert len(polys) == 100 and all(p.shape[1] == 2 and 5 <= len(p) <= 20 for p in polys)
    assert abs(np.mean([len(p) for p in polys]) - 14.4) < 0.2
    suite = U.load_default_train_data()
    assert len(suite.locations) == 100 and [list(map(list, p)) for p in suite.waypoint_suite] == [p.tolist() for p in polys]
    assert len(U.load_default_validation_data().locations) == 5
    ref = "/root/reference/torchdriveenv/data/training_cases.yml"
    if os.path.exists(ref):       # the packaged copy against the reference's file, loaded by the same loader
        want = U.load_waypoint_suite_data(ref)
        assert want.locations == suite.locations and want.waypoint_suite == suite.waypoint_suite
        assert want.car_sequence_suite == suite.car_sequence_suite
        assert [None if s is None else (s.agent_states, s.agent_attributes) for s in want.scenarios] == \
               [None if s is None else (s.agent_states, s.agent_attributes) for s in suite.scenarios]
    ss = S.training_mix(12, 5, seed=2)
    assert len(ss.maps) == 12 and len(ss.scenarios) == 12 and ss.max_agents() == 5
    for sc, p in zip(ss.scenarios, polys):
        assert np.array_equal(np.asarray(sc.waypoints, np.float32), p.astype(np.float32))
    packed = ss.pack(5)
    assert packed["scen_map"].tolist() == list(range(12))
    again = S.training_mix(12, 5, seed=2).pack(5)
    assert all(np.array_equal(packed[k], again[k]) for k in packed)      # seeded and deterministic
    extra = S.training_mix(102, 3, seed=1).scenarios[-1]                   # past the 100: synthetic, same statistics
    seg = np.hypot(*np.diff(np.asarray(extra.waypoints, np.float64), axis=0).T)
    assert 5 <= len(extra.waypoints) <= 20 and seg.min() > 12.8 and seg.max() < 15.1


def test_save_video_writes_the_frames(tmp_path):
    """helpers.save_video (reference helpers.py:7-36): list of B x 3 x H x W uint8 frames -> mp4v file at 10 fps."""
    import torch
    from torchdriveenv_b200.helpers import save_video, set_seeds
    frames = []
    for k in range(8):
        f = torch.zeros((1, 3, 64, 96), dtype=torch.uint8)
        f[0, 0, :, : 12 * (k + 1)] = 255
        frames.append(f)
    fn = str(tmp_path / "v.mp4")
    save_video(frames, fn)
    assert os.path.getsize(fn) > 200
    import cv2
    cap = cv2.VideoCapture(fn)
    assert int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)) == 96 and int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)) == 64
    assert int(round(cap.get(cv2.CAP_PROP_FPS))) == 10
    assert set_seeds(123) == 123


def test_vec_env_sb3_surface_without_an_engine():
    """The attribute / method part of SB3's VecEnv surface is plain host logic: exercised on an instance whose engine is
    never created (no GPU here)."""
    import torch
    from torchdriveenv_b200.gym_env import TorchDriveVecEnv
    v = object.__new__(TorchDriveVecEnv)
    v.num_envs, v.n_stack, v._seed = 5, 3, 11
    v._stack = torch.arange(5 * 9 * 64 * 64, dtype=torch.int64).remainder(251).to(torch.uint8).reshape(5, 9, 64, 64)
    assert v.get_attr("num_envs") == [5] * 5 and v.get_attr("n_stack", indices=[1, 3]) == [3, 3]
    v.set_attr("tag", "x")
    assert v.get_attr("tag", indices=2) == ["x"]
    assert v.env_is_wrapped(object) == [False] * 5
    assert v.seed(40) == [40, 41, 42, 43, 44] and v._seed == 40
    assert v.env_method("seed", indices=[0, 1]) == [[40, 41, 42, 43, 44]] * 2
    img = v.render()
    assert img.shape == (64, 64, 3) and np.array_equal(img[:, :, 0], v._stack[0, 6].numpy())
    assert len(v.get_images()) == 5 and v.get_images()[4].shape == (64, 64, 3)
    assert v.unwrapped is v and v.render_mode == "rgb_array"
