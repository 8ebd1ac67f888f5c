"""Step ordering, reward, termination, truncation, info and replay conventions of the oracle against
the reference's own source (torchdriveenv/gym_env.py, cited per test) re-derived in plain Python."""
import math

import numpy as np
import pytest

from torchdriveenv_b200 import scenarios as S
from torchdriveenv_b200._capi import INFO_COLUMNS as IC, default_config


def _straight_road(length=400.0, npcs=(), replay=None, lights=None):
    tris = []
    for k in range(int(length / 10)):
        x0, x1 = -20 + 10 * k, -10 + 10 * k
        tris += [[x0, -3.5, x1, -3.5, x1, 3.5, 1, 0], [x0, -3.5, x1, 3.5, x0, 3.5, 1, 0]]
    stop = np.zeros((0, 5), np.float32) if lights is None else np.asarray(lights[0], np.float32)
    ls = np.zeros((1, 0), np.uint8) if lights is None else np.asarray(lights[1], np.uint8)
    m = S.MapData(road_tris=np.asarray(tris, np.float32), stoplines=stop, light_states=ls)
    init = np.asarray([[0, 0, 0, 0]] + [list(n[:4]) for n in npcs], np.float32)
    attr = np.asarray([[5, 2, 0.9]] + [[5, 2, 2.0] for _ in npcs], np.float32)
    wp = np.asarray([[0, 0], [10, 0], [24, 0], [38, 0]], np.float32)
    rs = rm = None
    if replay is not None:
        rs, rm = replay
    sc = S.ScenarioData(0, wp, 0.0, init, attr, rs, rm)
    return S.ScenarioSet([m], [sc])


def _env(oracle, ss, A=None, **cfg):
    A = A or ss.max_agents()
    c = default_config(num_envs=1, max_agents=A, start_heading_sigma=0.0, start_speed_max=0.0, **cfg)
    env = oracle.OracleEnvSet(c, ss.pack(A))
    env.reset(seed=0)
    env.state[0, 0, :2] = [0.0, 0.0]
    return env


def test_reward_formula_matches_get_reward(oracle):
    """gym_env.py:396-411 evaluated in float64 Python next to the oracle's binary32."""
    env = _env(oracle, _straight_road())
    env.state[0, 0] = [0, 0, 0.1, 6.0]
    rng = np.random.default_rng(0)
    target_idx = 1
    wps = [[0, 0], [10, 0], [24, 0], [38, 0]]
    reached = 0
    for k in range(60):
        prev = env.state[0, 0].astype(np.float64).copy()
        a = np.array([[rng.uniform(-1, 1), rng.uniform(-0.05, 0.05)]], np.float32)
        _, r, term, trunc, info = env.step(a, render=False)
        cur = env.state[0, 0].astype(np.float64)
        d = math.dist(cur[:2], prev[:2])
        dist_r = 1.0 if d > 0.5 else 0.0
        psi_r = (1 - math.cos(cur[2] - prev[2])) * -25.0
        hit = target_idx < len(wps) and math.dist(cur[:2], wps[target_idx]) < 3
        want = (100.0 if hit else 0.0) + dist_r + psi_r
        assert abs(float(r[0]) - want) <= 1e-5 * max(1.0, abs(want))
        if hit:
            reached += 1; target_idx += 1
        assert info[0, IC["reached_waypoint_num"]] == reached
        assert abs(info[0, IC["psi_smoothness"]] - abs(prev[2] - cur[2]) / 0.1) < 1e-4
        assert abs(info[0, IC["speed_smoothness"]] - abs(prev[3] - cur[3]) / 0.1) < 1e-4
        assert abs(info[0, IC["psi_reward"]] - psi_r) < 1e-5 and info[0, IC["dist_reward"]] == dist_r
        assert env.env_vars[0, 2] == target_idx
        if term[0]:
            break
    assert reached >= 2


def test_stationary_ego_gets_zero_reward(oracle):
    env = _env(oracle, _straight_road())
    env.state[0, 0] = [5, 0, 0, 0]   # 5 m from waypoint 1: not reached
    _, r, term, trunc, info = env.step(np.zeros((1, 2), np.float32), render=False)
    assert r[0] == 0.0 and not term[0] and not trunc[0]


def test_waypoint_bonus_once_then_target_advances(oracle):
    env = _env(oracle, _straight_road())
    env.state[0, 0] = [8.0, 0, 0, 0]   # within 3 m of waypoint 1
    _, r, *_ = env.step(np.zeros((1, 2), np.float32), render=False)
    assert r[0] == 100.0 and env.env_vars[0, 2] == 2 and env.env_vars[0, 3] == 1
    _, r, *_ = env.step(np.zeros((1, 2), np.float32), render=False)
    assert r[0] == 0.0 and env.env_vars[0, 2] == 2          # next target is 14 m away
    env.env_vars[0, 2] = 4                                   # past the last waypoint: current_target = None (:380-383)
    env.state[0, 0] = [38.0, 0, 0, 0]
    _, r, *_ = env.step(np.zeros((1, 2), np.float32), render=False)
    assert r[0] == 0.0


def test_pure_yaw_penalty(oracle):
    env = _env(oracle, _straight_road())
    env.state[0, 0] = [5, 0, 0, 2.0]
    _, r, _, _, info = env.step(np.array([[0.0, 0.3]], np.float32), render=False)
    dpsi = float(env.state[0, 0, 2])
    assert abs(dpsi - (2.0 / 0.9) * math.sin(0.3) * 0.1) < 1e-6
    assert abs(float(r[0]) - (-25 * (1 - math.cos(dpsi)))) < 1e-5   # moved 0.2 m < cutoff: no distance bonus


def test_truncation_at_max_steps_and_is_success(oracle):
    env = _env(oracle, _straight_road(), max_environment_steps=5)
    for k in range(5):
        _, r, term, trunc, info = env.step(np.zeros((1, 2), np.float32), render=False)
        assert bool(trunc[0]) == (k == 4) and bool(info[0, IC["is_success"]]) == (k == 4)   # :134-135, :430
    assert env.env_vars[0, 1] == 5


def test_termination_on_each_infraction(oracle):
    # offroad
    env = _env(oracle, _straight_road())
    env.state[0, 0] = [5, 3.0, 0, 0]      # corners at y = 4.0: 0.5 m off, equal to the threshold -> not offroad
    _, _, term, _, info = env.step(np.zeros((1, 2), np.float32), render=False)
    assert not term[0] and info[0, IC["offroad"]] == 0
    env.state[0, 0] = [5, 3.2, 0, 0]
    _, _, term, _, info = env.step(np.zeros((1, 2), np.float32), render=False)
    assert term[0] and abs(info[0, IC["offroad"]] - 2 * 0.2) < 1e-4
    # terminated_at_infraction = False (:416-417)
    env = _env(oracle, _straight_road(), terminated_at_infraction=0)
    env.state[0, 0] = [5, 3.2, 0, 0]
    _, _, term, _, info = env.step(np.zeros((1, 2), np.float32), render=False)
    assert not term[0] and info[0, IC["offroad"]] > 0
    # collision with a parked NPC
    env = _env(oracle, _straight_road(npcs=[[12, 0, 0, 0]]))
    env.state[0, 0] = [6.5, 0, 0, 0]
    _, _, term, _, info = env.step(np.zeros((1, 2), np.float32), render=False)
    assert not term[0] and info[0, IC["collision"]] == 0      # 0.5 m gap
    env.state[0, 0] = [7.1, 0, 0, 0]
    _, _, term, _, info = env.step(np.zeros((1, 2), np.float32), render=False)
    assert term[0] and info[0, IC["collision"]] == 1
    assert env.infractions[0, 1, 0] == 1                        # the NPC sees the same collision


def test_red_light_violation_uses_rear_strip(oracle):
    stop = [[20.0, 0.0, 0.8, 3.5, 0.0]]
    sched = np.array([[2], [2], [0], [0]], np.uint8)            # red, red, green, green
    env = _env(oracle, _straight_road(lights=(stop, sched)))
    env.env_vars[0, 4] = 0
    env.state[0, 0] = [19.0, 0, 0, 0]   # front over the line, rear strip (x in [16.5, 17]) not yet
    _, _, term, _, info = env.step(np.zeros((1, 2), np.float32), render=False)
    assert not term[0] and info[0, IC["traffic_light_violation"]] == 0
    env = _env(oracle, _straight_road(lights=(stop, sched)))
    env.env_vars[0, 4] = 0
    env.state[0, 0] = [22.3, 0, 0, 0]   # rear strip x in [19.8, 20.3] overlaps the line [19.6, 20.4], light red at t=1
    _, _, term, _, info = env.step(np.zeros((1, 2), np.float32), render=False)
    assert term[0] and info[0, IC["traffic_light_violation"]] == 1
    env = _env(oracle, _straight_road(lights=(stop, sched)))
    env.env_vars[0, 4] = 1              # phase 1: t = 2 -> green
    env.state[0, 0] = [22.3, 0, 0, 0]
    _, _, term, _, info = env.step(np.zeros((1, 2), np.float32), render=False)
    assert not term[0] and info[0, IC["traffic_light_violation"]] == 0


def test_wrong_way(oracle):
    env = _env(oracle, _straight_road())
    env.state[0, 0] = [5, 0, math.pi - 0.2, 0]
    inf = env.compute_infractions()
    assert abs(inf[0, 0, 3] - math.cos(0.2)) < 1e-5
    env.state[0, 0] = [5, 0, 1.0, 0]
    assert env.compute_infractions()[0, 0, 3] == 0.0


def test_replay_and_constant_velocity_npcs(oracle):
    T = 6
    rs = np.zeros((T, 3, 4), np.float32); rm = np.zeros((T, 3), np.uint8)
    rs[:, 1] = [[30 + t, 0, 0, 10] for t in range(T)]; rm[:, 1] = 1        # replayed
    ss = _straight_road(npcs=[[30, 0, 0, 10], [60, 0, 0, 5]], replay=(rs, rm))
    env = _env(oracle, ss)
    np.testing.assert_array_equal(env.state[0, 1], rs[0, 1])               # reset state = replay[0]
    for t in range(1, 9):
        env.step(np.zeros((1, 2), np.float32), render=False)
        if t < T:
            np.testing.assert_array_equal(env.state[0, 1], rs[t, 1])       # overwritten from the log (:275-294)
        else:
            assert abs(env.state[0, 1, 0] - (rs[T - 1, 1, 0] + (t - T + 1) * 1.0)) < 1e-4   # then constant velocity
        assert abs(env.state[0, 2, 0] - (60 + 0.5 * t)) < 1e-4 and env.state[0, 2, 3] == 5.0  # never-replayed NPC


def test_absent_agents_do_not_move_or_collide(oracle):
    ss = _straight_road(npcs=[[3, 0, 0, 4]])
    env = _env(oracle, ss, A=4)
    assert list(env.attr[0, :, 3]) == [1, 1, 0, 0]
    before = env.state[0, 2:].copy()
    _, _, term, _, info = env.step(np.zeros((1, 2), np.float32), render=False)
    np.testing.assert_array_equal(env.state[0, 2:], before)
    assert info[0, IC["collision"]] == 1      # only the present NPC counts (absent slots sit at the origin too)


def test_reset_sampling_rules(oracle):
    """set_start_pos :351-367: start on the first waypoint segment, speed in [0, 10), heading = lane
    direction + noise; target index 1 (:325); counters zero (:338-339)."""
    ss = S.traffic_lights(8)
    E = 2000
    c = default_config(num_envs=E, max_agents=8, randomize_ego_attributes=1)
    env = oracle.OracleEnvSet(c, ss.pack(8))
    env.reset(seed=11)
    wp = ss.scenarios[0].waypoints.astype(np.float64)
    p = env.state[:, 0, :2].astype(np.float64)
    u = (p - wp[0]) @ (wp[1] - wp[0]) / np.dot(wp[1] - wp[0], wp[1] - wp[0])
    assert (u >= -1e-6).all() and (u <= 1 + 1e-6).all() and 0.45 < u.mean() < 0.55
    resid = p - (wp[0] + u[:, None] * (wp[1] - wp[0]))
    assert np.abs(resid).max() < 1e-3
    v = env.state[:, 0, 3]
    assert v.min() >= 0 and v.max() < 10 and 4.5 < v.mean() < 5.5
    dpsi = env.state[:, 0, 2] - ss.scenarios[0].start_heading
    assert abs(dpsi.mean()) < 0.01 and 0.09 < dpsi.std() < 0.11
    assert (env.env_vars[:, 1] == 0).all() and (env.env_vars[:, 2] == 1).all() and (env.env_vars[:, 3] == 0).all()
    at = env.attr[:, 0]
    assert at[:, 0].min() >= 4.8 and at[:, 0].max() <= 5.5 and at[:, 1].min() >= 1.8 and at[:, 1].max() <= 2.2
    assert at[:, 2].min() >= 0.82 and at[:, 2].max() <= 0.97
    # a different episode index draws a different start; the same seed reproduces
    first = env.state.copy()
    env.reset(seed=11)
    assert not np.array_equal(first, env.state)
    env2 = oracle.OracleEnvSet(c, ss.pack(8)); env2.reset(seed=11)
    np.testing.assert_array_equal(first, env2.state)


def test_sharded_envs_reproduce_the_unsharded_run(oracle):
    ss = S.roundabout(8)
    packed = ss.pack(8)
    full = oracle.OracleEnvSet(default_config(num_envs=12, max_agents=8, auto_reset=1), packed)
    lo = oracle.OracleEnvSet(default_config(num_envs=5, max_agents=8, auto_reset=1, env_index_offset=0), packed)
    hi = oracle.OracleEnvSet(default_config(num_envs=7, max_agents=8, auto_reset=1, env_index_offset=5), packed)
    for e in (full, lo, hi):
        e.reset(seed=4)
    rng = np.random.default_rng(4)
    for _ in range(40):
        a = np.stack([rng.uniform(-1, 1, 12), rng.uniform(-0.3, 0.3, 12)], 1).astype(np.float32)
        full.step(a, render=False); lo.step(a[:5], render=False); hi.step(a[5:], render=False)
    np.testing.assert_array_equal(full.state, np.concatenate([lo.state, hi.state]))
    np.testing.assert_allclose(full.stats, lo.stats + hi.stats, rtol=1e-12)
    assert full.stats[0] > 0


def test_auto_reset_flags_and_statistics(oracle):
    ss = S.traffic_lights(8)
    E = 64
    env = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=8, auto_reset=1, max_environment_steps=20), ss.pack(8))
    env.reset(seed=2)
    rng = np.random.default_rng(2)
    episodes = 0
    ret = np.zeros(E); ret_sum = 0.0
    for _ in range(60):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1)
        _, r, term, trunc, info = env.step(a, render=False)
        done = (term | trunc).astype(bool)
        ret += r
        np.testing.assert_array_equal(info[:, IC["did_reset"]] != 0, done)
        assert (env.env_vars[done, 1] == 0).all()                      # step counter restarted
        assert np.allclose(info[:, IC["episode_return"]], ret, rtol=1e-5, atol=1e-4)
        ret_sum += ret[done].sum(); ret[done] = 0
        episodes += int(done.sum())
    assert episodes > E and env.stats[0] == episodes and env.stats[8] == 60 * E
    assert abs(env.stats[1] - ret_sum) < 1e-2
