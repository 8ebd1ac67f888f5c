"""The SB3-shaped `infos` of TorchDriveVecEnv (gym_env.VecInfos) against its consumer in the reference: every per-env dict
of every step of an auto-reset run goes through the reference's own, unmodified EvalNTimestepsCallback._calc_metrics
(examples/rl_training.py:39-67; imported where the checkout exists, restated line by line where it does not - the GPU
box), and the callback's counters must equal the episode-statistics vector.  The arrays come from the oracle here (same
layout as the C ABI's info rows); tests/test_gpu_env_api.py checks that the product VecEnv hands out the same dicts."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_RL = "/root/reference/examples/rl_training.py"


class RestatedCallback:
    """EvalNTimestepsCallback._calc_metrics, examples/rl_training.py:39-67, restated (used only without the checkout)."""

    def __init__(self):
        self.episode_num = self.offroad_num = self.collision_num = self.traffic_light_violation_num = self.success_num = 0
        self.reached_waypoint_nums, self.psi_smoothness, self.speed_smoothness = [], [], []
        self.psi_smoothness_for_single_episode, self.speed_smoothness_for_single_episode = [], []

    def _calc_metrics(self, locals_, globals_):
        info = locals_["info"]
        if "psi_smoothness" not in info:                                                      # :47-48
            return
        self.psi_smoothness_for_single_episode.append(info["psi_smoothness"])                # :49
        self.speed_smoothness_for_single_episode.append(info["speed_smoothness"])            # :50
        if (info["offroad"] > 0) or (info["collision"] > 0) or (info["traffic_light_violation"] > 0) or (info["is_success"]):  # :51-52
            self.episode_num += 1
            self.offroad_num += int(info["offroad"] > 0)
            self.collision_num += int(info["collision"] > 0)
            self.traffic_light_violation_num += int(info["traffic_light_violation"] > 0)
            self.success_num += int(bool(info["is_success"]))
            self.reached_waypoint_nums.append(info["reached_waypoint_num"])


def make_callback():
    if os.path.exists(REFERENCE_RL):
        spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(HERE, "golden", "make_reference_golden.py"))
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
        rl = m.import_reference_eval_callback()
        cb = rl.EvalNTimestepsCallback(eval_env=None, n_steps=1, eval_n_episodes=1)
        cb.episode_num = cb.offroad_num = cb.collision_num = cb.traffic_light_violation_num = cb.success_num = 0
        cb.reached_waypoint_nums, cb.psi_smoothness, cb.speed_smoothness = [], [], []
        cb.psi_smoothness_for_single_episode, cb.speed_smoothness_for_single_episode = [], []
        return cb, "reference"
    return RestatedCallback(), "restated"


def feed(cb, infos, open_eps):
    """One vectorised step: env i's dict goes to the callback; a finished env starts a new per-episode accumulator
    (what _evaluate :83-95 does between episodes)."""
    for i in range(len(infos)):
        d = infos[i]
        cb.psi_smoothness_for_single_episode, cb.speed_smoothness_for_single_episode = open_eps[i]
        cb._calc_metrics({"info": d}, {})
        if "episode" in d:
            open_eps[i] = ([], [])


def test_vec_infos_drive_the_reference_metrics(oracle):
    from torchdriveenv_b200 import scenarios as S
    from torchdriveenv_b200._capi import STAT_NAMES, default_config
    from torchdriveenv_b200.gym_env import VecInfos
    E, steps = 32, 150
    ss = S.validation_mix(8)
    A = ss.max_agents()
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1, max_environment_steps=40), ss.pack(A))
    orc.reset(seed=17)
    rng = np.random.default_rng(17)
    cb, kind = make_callback()
    open_eps = [([], []) for _ in range(E)]
    n_done = 0
    for _ in range(steps):
        a = np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32)
        _, rew, term, trunc, info = orc.step(a, render=False)
        infos = VecInfos(info, term.astype(bool), trunc.astype(bool))
        assert len(infos) == E and isinstance(infos[0], dict)
        for i in np.nonzero(term | trunc)[0]:
            d = infos[int(i)]
            assert d["TimeLimit.truncated"] == bool(trunc[i] and not term[i])
            assert d["episode"]["l"] == int(info[i, 11]) and d["episode"]["r"] == float(info[i, 10])
            n_done += 1
        feed(cb, infos, open_eps)
    stats = dict(zip(STAT_NAMES, orc.stats))
    assert n_done == int(stats["episodes"]) and n_done > 30
    got = dict(episodes=cb.episode_num, offroad=cb.offroad_num, collision=cb.collision_num,
               traffic_light_violation=cb.traffic_light_violation_num, success=cb.success_num,
               reached_waypoints=float(sum(cb.reached_waypoint_nums)))
    for k, v in got.items():
        assert float(v) == float(stats[k]), (kind, k, v, stats[k])
