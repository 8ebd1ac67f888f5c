"""The reference-shaped Python API (gym.Env / SingleAgentWrapper / SimulatorInterface / VecEnv) on top
of the CUDA engine, checked against the oracle."""
import numpy as np
import pytest
import torch

from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200._capi import default_config

pytestmark = pytest.mark.gpu


def _suite():
    from torchdriveenv_b200.gym_env import Scenario, WaypointSuite
    return WaypointSuite(locations=["Town07", "Town01"],
                         waypoint_suite=[S.VALIDATION_POLYLINES["three_way"], S.VALIDATION_POLYLINES["traffic_lights"]],
                         car_sequence_suite=[{}, {}],
                         scenarios=[Scenario(agent_states=S.THREE_WAY_NPCS["states"], agent_attributes=S.THREE_WAY_NPCS["attributes"],
                                             recurrent_states=[[0.0] * 4] * 2), None])


def test_single_env_api_matches_reference_shapes(oracle):
    from torchdriveenv_b200.gym_env import EnvConfig, SingleAgentWrapper, WaypointSuiteEnv
    cfg = EnvConfig(seed=3, device="cuda:0")
    env = SingleAgentWrapper(WaypointSuiteEnv(cfg, _suite(), n_background=4))
    assert env.action_space.shape == (2,) and env.observation_space.shape == (3, 64, 64)
    np.testing.assert_allclose(env.action_space.low, [-1.0, -0.3]); np.testing.assert_allclose(env.action_space.high, [1.0, 0.3])
    obs, info = env.reset()
    assert obs.shape == (3, 64, 64) and obs.dtype == np.uint8 and info == {}
    inner = env.env
    orc = oracle.OracleEnvSet(default_config(num_envs=1, max_agents=inner.engine.A), inner.engine.packed)
    orc.reset(seed=3)   # the simulator is reset once when it is built ...
    orc.reset(seed=3)   # ... and once more by env.reset(): second episode of the same seed
    assert np.array_equal(obs, orc.render()[0])
    total = 0.0
    for k in range(40):
        a = np.array([1.0, 0.0], np.float32)            # the notebook's constant action [1, 0]
        obs, reward, terminated, truncated, info = env.step(a)
        oobs, orr, ote, otr, oinfo = orc.step(a[None])
        assert obs.shape == (3, 64, 64) and isinstance(reward, float) and isinstance(terminated, bool) and isinstance(truncated, bool)
        assert np.array_equal(obs, oobs[0]) and reward == float(orr[0]) and terminated == bool(ote[0]) and truncated == bool(otr[0])
        for key in ("offroad", "collision", "traffic_light_violation", "is_success", "reached_waypoint_num", "psi_smoothness",
                    "psi_reward", "dist_reward", "speed_smoothness"):
            assert key in info
        assert float(info["offroad"]) == float(oinfo[0, 0]) and info["reached_waypoint_num"] == int(oinfo[0, 4])
        total += reward
        if terminated or truncated:
            break
    frame = env.render()
    assert frame.shape == (64, 64, 3)
    env.close()


def test_simulator_interface_surface(oracle):
    from torchdriveenv_b200.simulator import BatchedSimulator
    ss = S.traffic_lights(12)
    sim = BatchedSimulator(ss, num_envs=32, device="cuda:0", seed=5)
    orc = oracle.OracleEnvSet(default_config(num_envs=32, max_agents=12), sim.engine.packed)
    orc.reset(seed=5)
    assert sim.get_state().shape == (32, 1, 4)                       # NPCs are hidden, as IAIWrapper hides them
    rng = np.random.default_rng(5)
    for _ in range(10):
        a = np.stack([rng.uniform(-1, 1, 32), rng.uniform(-0.3, 0.3, 32)], 1).astype(np.float32)
        sim.step(torch.from_numpy(a).view(32, 1, 2))
        orc.step(a, render=False, phases=3)
    assert np.array_equal(sim.get_state().cpu().numpy(), orc.state[:, :1])
    assert np.array_equal(sim.compute_offroad().cpu().numpy(), orc.infractions[:, :1, 1])
    assert np.array_equal(sim.compute_collision().cpu().numpy(), orc.infractions[:, :1, 0])
    assert np.array_equal(sim.compute_traffic_lights_violations().cpu().numpy(), orc.infractions[:, :1, 2])
    assert np.array_equal(sim.compute_wrong_way().cpu().numpy(), orc.infractions[:, :1, 3])
    bv = sim.render_egocentric()
    assert bv.shape == (32, 1, 3, 64, 64) and bv.dtype == torch.uint8
    assert np.array_equal(bv[:, 0].cpu().numpy(), orc.render())
    twin = sim.copy()
    assert torch.equal(twin.get_state(), sim.get_state()) and torch.equal(twin.render_egocentric(), bv)
    twin.step(torch.zeros(32, 1, 2))
    assert not torch.equal(twin.get_state(), sim.get_state())        # independent after the copy
    assert sim.to("cuda:0") is sim
    with pytest.raises(RuntimeError, match="no CPU path"):
        sim.to("cpu")
    full = BatchedSimulator(ss, num_envs=4, device="cuda:0", expose_npcs=True)
    assert full.get_state().shape == (4, 12, 4) and full.compute_collision().shape == (4, 12)


@pytest.mark.parametrize("frame_ring", [None, 0, 3])
def test_vec_env_frame_stack_and_auto_reset(oracle, frame_ring):
    """frame_ring=None: the default ring of n_stack + 1 stacked observations (no frame is moved; the tensor a step returns
    survives the next step); 0: one tensor shifted in place by the kernel; 3: a ring without a spare slot."""
    from torchdriveenv_b200.gym_env import EnvConfig, TorchDriveVecEnv
    E = 48
    cfg = EnvConfig(seed=8, device="cuda:0", max_environment_steps=15)
    venv = TorchDriveVecEnv(cfg, S.traffic_lights(8), num_envs=E, n_stack=3, frame_ring=frame_ring)
    last_obs, last_want = None, None
    obs = venv.reset()
    assert obs.shape == (E, 9, 64, 64) and obs.dtype == torch.uint8
    assert bool((obs[:, :6] == 0).all()) and bool((obs[:, 6:] != 0).any())
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=8, auto_reset=1, max_environment_steps=15), venv.engine.packed)
    orc.reset(seed=8)
    rng = np.random.default_rng(8)
    frames = [orc.render()]
    age = np.zeros(E, np.int64)      # frames since the env (re)started: slot s holds a frame only if age >= 2 - s
    n_done = 0
    for k in range(25):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, rew, dones, infos = venv.step(a)
        oobs, orr, ote, otr, oinfo = orc.step(a)
        assert np.array_equal(obs[:, 6:].cpu().numpy(), oobs) and np.array_equal(rew.cpu().numpy(), orr)
        d = (ote | otr).astype(bool)
        assert np.array_equal(dones.cpu().numpy(), d)
        # channels-first stack: previous frame in the middle slot unless the env just restarted
        prev = frames[-1]
        keep = ~d
        assert np.array_equal(obs[keep][:, 3:6].cpu().numpy(), prev[keep])
        assert bool((obs[d][:, :6] == 0).all())
        frames.append(oobs)
        age = np.where(d, 0, age + 1)
        # the whole stack against VecFrameStack semantics (oldest frame first, zeros before the restart)
        want = np.zeros((E, 9, 64, 64), np.uint8)
        for slot in range(3):
            back = 2 - slot
            if back < len(frames):
                have = age >= back
                want[have, 3 * slot:3 * slot + 3] = frames[-1 - back][have]
        assert np.array_equal(obs.cpu().numpy(), want), f"step {k}: fused frame stack"
        if frame_ring is None and last_obs is not None:   # the observation the policy acted on is still there after the step
            assert np.array_equal(last_obs.cpu().numpy(), last_want), f"step {k}: the previous observation was overwritten"
        last_obs, last_want = obs, want
        n_done += int(d.sum())
        assert set(infos.columns) >= {"offroad", "collision", "traffic_light_violation", "is_success", "terminated", "truncated"}
        # SB3's contract: a sequence of E dicts; finished envs carry Monitor's episode record and TimeLimit.truncated
        assert len(infos) == E
        for e in range(E):
            row = infos[e]
            assert row["offroad"] == float(oinfo[e, 0]) and row["is_success"] == bool(otr[e])
            assert ("episode" in row) == bool(d[e])
            if d[e]:
                assert row["episode"]["l"] == int(oinfo[e, 11]) and row["episode"]["r"] == float(oinfo[e, 10])
                assert row["TimeLimit.truncated"] == bool(otr[e] and not ote[e])
    assert n_done > 0
    stats = venv.episode_statistics()
    assert stats["episodes"] == n_done and stats["steps"] == 25 * E
    venv.close()


def test_background_traffic_agents_drive_at_constant_velocity(oracle):
    """Agents taken from a background-traffic file (gym_env.py:200-233 schema) have no replay: the step
    kernel moves them with the bicycle model and zero action, exactly as the oracle does."""
    from torchdriveenv_b200.gym_env import EnvConfig, TorchDriveVecEnv, WaypointSuite, scenario_set_from_suite
    rng = np.random.default_rng(5)
    poly = S.VALIDATION_POLYLINES["traffic_lights"]
    n = 40
    bt = dict(location="carla:Town01", agent_density=10, random_seed=1,
              states=np.stack([rng.uniform(60, 220, n), rng.uniform(-20, 100, n), rng.uniform(-3, 3, n), rng.uniform(0, 8, n)], 1).astype(np.float32),
              attributes=np.stack([rng.uniform(4.2, 5.2, n), rng.uniform(1.8, 2.1, n), rng.uniform(1.0, 1.6, n)], 1).astype(np.float32))
    suite = WaypointSuite(locations=["Town01"], waypoint_suite=[poly], car_sequence_suite=[None], scenarios=[None])
    cfg = EnvConfig(seed=5, device="cuda:0", max_environment_steps=30)
    ss = scenario_set_from_suite(cfg, suite, background_traffic=bt)
    far = np.hypot(bt["states"][:, 0] - poly[0][0], bt["states"][:, 1] - poly[0][1]) > 100
    A = ss.max_agents()
    assert A == 1 + int(far.sum()) and A > 4
    E = 16
    venv = TorchDriveVecEnv(cfg, ss, num_envs=E, n_stack=1)
    obs = venv.reset()
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1, max_environment_steps=30), venv.engine.packed)
    orc.reset(seed=5)
    assert np.array_equal(obs.cpu().numpy(), orc.render())
    for k in range(12):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, rew, dones, infos = venv.step(a)
        oobs, orr, ote, otr, oinfo = orc.step(a)
        assert np.array_equal(obs.cpu().numpy(), oobs) and np.array_equal(rew.cpu().numpy(), orr)
        assert np.array_equal(venv.engine.get_state().cpu().numpy(), orc.state)
    st = venv.engine.get_state().cpu().numpy()
    keep = ~(ote | otr).astype(bool)          # envs that were not just re-initialised
    moved = np.hypot(st[keep][:, 1:, 0] - ss.scenarios[0].agent_init[1:, 0], st[keep][:, 1:, 1] - ss.scenarios[0].agent_init[1:, 1])
    assert keep.any() and (moved[:, ss.scenarios[0].agent_init[1:, 3] > 0.5] > 0).all()
    venv.close()


@pytest.mark.parametrize("frame_copy", ["scatter", "shift"])
def test_rollout_collector_fills_the_buffer_like_vec_frame_stack(oracle, frame_copy):
    """Config C5 in small: RolloutCollector on the training-scenario mix.  Every slot of the GPU rollout buffer
    must hold what SubprocVecEnv + VecFrameStack + RolloutBuffer.add would have stored (examples/rl_training.py
    :159-160): oracle frames stacked oldest first, zeros before a restart, rewards / flags / episode starts."""
    from torchdriveenv_b200.engine import Engine
    from torchdriveenv_b200.rollout import RolloutCollector
    E, A, T, NS = 40, 6, 9, 3
    ss = S.training_mix(7, A, seed=3)
    eng = Engine(ss, E, A, device="cuda:0", auto_reset=1, max_environment_steps=12)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1, max_environment_steps=12), eng.packed)
    orc.reset(seed=11)
    col = RolloutCollector(eng, T, n_stack=NS, seed=11, with_info=True, frame_copy=frame_copy)
    rng = np.random.default_rng(11)
    acts = np.stack([rng.uniform(-1, 1, (3 * T, E)), rng.uniform(-0.3, 0.3, (3 * T, E))], -1).astype(np.float32)
    k = [0]

    def policy(obs):
        assert obs.shape == (E, 3 * NS, 64, 64) and obs.dtype == torch.uint8
        a = torch.from_numpy(acts[k[0]]).to(obs.device)
        k[0] += 1
        return a, a[:, 0] * 2.0, a[:, 1] - 1.0         # (actions, values, log_probs)

    frames = [orc.render()]
    age = np.zeros(E, np.int64)

    def want_stack():
        w = np.zeros((E, 3 * NS, 64, 64), np.uint8)
        for slot in range(NS):
            back = NS - 1 - slot
            if back < len(frames):
                have = age >= back
                w[have, 3 * slot:3 * slot + 3] = frames[-1 - back][have]
        return w

    n_done = 0
    for r in range(3):
        b = col.collect(policy)
        torch.cuda.synchronize()
        obs = b.observations.cpu().numpy()
        assert np.array_equal(obs[0], want_stack()), f"rollout {r}: slot 0 carries the last observation over"
        starts = b.episode_starts.cpu().numpy()
        assert (starts[0] == (1 if r == 0 else d_last)).all()
        for t in range(T):
            a = acts[r * T + t]
            oobs, orr, ote, otr, oinfo = orc.step(a)
            d = (ote | otr).astype(bool)
            frames.append(oobs)
            age = np.where(d, 0, age + 1)
            assert np.array_equal(obs[t + 1], want_stack()), f"rollout {r} step {t}: buffer slot"
            assert np.array_equal(b.actions[t].cpu().numpy(), a)
            assert np.array_equal(b.rewards[t].cpu().numpy(), orr)
            assert np.array_equal(b.terminated[t].cpu().numpy(), ote) and np.array_equal(b.truncated[t].cpu().numpy(), otr)
            assert np.array_equal(starts[t + 1].astype(bool), d)
            assert np.array_equal(b.infos[t].cpu().numpy(), oinfo)
            assert np.array_equal(b.values[t].cpu().numpy(), a[:, 0] * 2.0) and np.array_equal(b.log_probs[t].cpu().numpy(), a[:, 1] - 1.0)
            n_done += int(d.sum())
            d_last = d.astype(np.uint8)
    assert n_done > 0 and col.num_timesteps == 3 * T * E
    assert col.episode_statistics()["episodes"] == n_done
    # GAE against a plain float64 loop
    last_v = torch.linspace(-1, 1, E, device="cuda:0")
    ret, adv = col.returns_and_advantages(last_v, gamma=0.99, gae_lambda=0.95)
    rw, va, st = b.rewards.cpu().numpy().astype(np.float64), b.values.cpu().numpy().astype(np.float64), b.episode_starts.cpu().numpy()
    want = np.zeros((T, E)); last = np.zeros(E); nxt = last_v.cpu().numpy().astype(np.float64)
    for t in reversed(range(T)):
        nt = 1.0 - st[t + 1]
        delta = rw[t] + 0.99 * nxt * nt - va[t]
        last = delta + 0.99 * 0.95 * nt * last
        want[t] = last; nxt = va[t]
    assert np.allclose(adv.cpu().numpy(), want, rtol=1e-4, atol=1e-3) and np.allclose(ret.cpu().numpy(), want + va, rtol=1e-4, atol=1e-3)
    # overlapping but different slots are refused, n_stack = 1 is refused
    from torchdriveenv_b200._capi import TdeError
    flat = torch.zeros(E * 9 * 4096 + 4096, dtype=torch.uint8, device="cuda:0")
    s0, s1 = flat[:E * 9 * 4096].view(E, 9, 64, 64), flat[4096:].view(E, 9, 64, 64)
    with pytest.raises(TdeError, match="overlap"):
        eng.step_rollout(torch.zeros(E, 2), s0, s1, 3)
    eng.close()


@pytest.mark.parametrize("n_stack", [1, 3])
def test_vec_env_terminal_observation(oracle, n_stack):
    """SB3 VecEnv contract (the caller, examples/rl_training.py:159): when an episode ends, infos carry the last
    observation of that episode (`terminal_observation`) and `obs` is the first one of the next.  Everything
    else must be what the in-kernel auto-reset path returns."""
    from torchdriveenv_b200.gym_env import EnvConfig, TorchDriveVecEnv
    E, A = 40, 8
    cfg = EnvConfig(seed=4, device="cuda:0", max_environment_steps=14)
    venv = TorchDriveVecEnv(cfg, S.traffic_lights(A), num_envs=E, n_stack=n_stack, terminal_observation=True)
    twin = TorchDriveVecEnv(cfg, S.traffic_lights(A), num_envs=E, n_stack=n_stack)          # in-kernel auto-reset
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1, max_environment_steps=14), venv.engine.packed)
    term_buf = np.zeros((E, 3, 64, 64), np.uint8)
    orc.set_terminal_buffer(term_buf)
    orc.reset(seed=4)
    obs = venv.reset(); obs2 = twin.reset()
    assert torch.equal(obs, obs2)
    rng = np.random.default_rng(4)
    n_done = 0
    prev_stack = obs.cpu().numpy().copy()
    for k in range(30):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        obs, rew, dones, infos = venv.step(a)
        obs2, rew2, dones2, infos2 = twin.step(a)
        oobs, orr, ote, otr, oinfo = orc.step(a)
        d = (ote | otr).astype(bool)
        assert torch.equal(obs, obs2) and torch.equal(rew, rew2) and torch.equal(dones, dones2)
        assert torch.equal(venv.engine.get_state(), twin.engine.get_state())
        assert torch.equal(venv.engine.get_env_vars(), twin.engine.get_env_vars())
        for key, col in infos2.columns.items():
            assert torch.equal(infos.columns[key], col), key
        assert np.array_equal(dones.cpu().numpy(), d) and np.array_equal(obs[:, -3:].cpu().numpy(), oobs)
        tobs = infos.columns["terminal_observation"].cpu().numpy()
        for e in np.nonzero(d)[0][:3]:      # the per-env dict of a finished env carries its terminal observation
            assert np.array_equal(infos[int(e)]["terminal_observation"], tobs[e])
        assert all("terminal_observation" not in infos[int(e)] for e in np.nonzero(~d)[0][:3])
        assert tobs.shape == (E, 3 * n_stack, 64, 64)
        if d.any():
            assert np.array_equal(tobs[d][:, -3:], term_buf[d]), f"step {k}: terminal frame"
            assert not np.array_equal(tobs[d][:, -3:], oobs[d])            # it is not the frame of the new episode
            if n_stack > 1:   # the older slots of the terminal stack are the previous stack shifted by one frame
                assert np.array_equal(tobs[d][:, :-3], prev_stack[d][:, 3:])
                assert bool((obs[torch.from_numpy(d).cuda()][:, :-3] == 0).all())
        n_done += int(d.sum())
        prev_stack = obs.cpu().numpy().copy()
    assert n_done > 5
    assert venv.episode_statistics() == twin.episode_statistics()
    venv.close(); twin.close()


def test_recording_view_matches_the_oracle(oracle):
    """tde_render_view (BirdviewRecordingWrapper's frame, gym_env.py:52-53,295-297) against orc_render_view: exact,
    at several resolutions, cameras and zooms, on a multi-map set (maps smaller than the scratch's largest)."""
    from torchdriveenv_b200.engine import Engine
    from torchdriveenv_b200._capi import default_config
    ss = S.validation_mix(8); A = 10; E = 10
    eng = Engine(ss, E, A, device="cuda:0", auto_reset=1)
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1), eng.packed)
    eng.reset(seed=9); orc.reset(seed=9)
    rng = np.random.default_rng(9)
    for _ in range(4):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        eng.step(torch.from_numpy(a).cuda()); orc.step(a)
    st = orc.state
    cases = [(0, float(st[0, 0, 0]), float(st[0, 0, 1]), float(st[0, 0, 2]), 35.0, 64, 64),
             (3, float(st[3, 0, 0]), float(st[3, 0, 1]), 0.0, 120.0, 256, 192),
             (5, float(st[5, 0, 0]) + 7.5, float(st[5, 0, 1]) - 3.25, 1.1, 60.0, 333, 127),
             (7, 0.0, 0.0, -0.4, 500.0, 1024, 1024),
             (9, float(st[9, 0, 0]), float(st[9, 0, 1]), 2.0, 8.0, 512, 512)]
    for (e, cx, cy, psi, fov, W, H) in cases:
        got = eng.render_view(e, (cx, cy), psi, fov, (W, H)).cpu().numpy()
        want = orc.render_view(e, cx, cy, psi, fov, W, H)
        assert got.shape == (3, H, W)
        assert np.array_equal(got, want), f"case env {e} {W}x{H}: {(got != want).mean()}"
    # the observation settings reproduce the observation
    assert np.array_equal(eng.render_view(0, (cases[0][1], cases[0][2]), cases[0][3], 35.0, (64, 64)).cpu().numpy(), eng.render()[0].cpu().numpy())
    with pytest.raises(Exception, match="TDE_E_INVAL"):
        eng.render_view(E, (0.0, 0.0), 0.0, 35.0, (64, 64))
    with pytest.raises(Exception, match="TDE_E_INVAL"):
        eng.render_view(0, (0.0, 0.0), 0.0, 35.0, (5000, 64))
    eng.close()


def test_video_render_mode_records_and_saves(tmp_path):
    """render_mode='video' (gym_env.py:295-297,170-177): one video_res x video_res frame after reset and after every
    step, written as mp4 by close()."""
    from torchdriveenv_b200 import gym_env as G
    fn = str(tmp_path / "episode.mp4")
    cfg = G.EnvConfig(seed=4, device="cuda:0", render_mode="video", video_filename=fn, video_res=256, video_fov=300)
    env = G.SingleAgentWrapper(G.WaypointSuiteEnv(cfg, S.three_way(6)))
    env.reset()
    for _ in range(5):
        env.step(np.array([0.5, 0.0], np.float32))
    bvs = env.env.simulator.get_birdviews()
    assert len(bvs) == 6 and tuple(bvs[0].shape) == (1, 3, 256, 256) and bvs[0].dtype == torch.uint8 and not bvs[0].is_cuda
    assert not torch.equal(bvs[0], bvs[-1])
    with pytest.raises(NotImplementedError):
        env.render()                      # the reference's render() only serves 'rgb_array' (gym_env.py:152-157)
    env.close()
    import os
    assert os.path.getsize(fn) > 1000
