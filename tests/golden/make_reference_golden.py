"""Golden vectors produced by the REFERENCE'S OWN CODE for the rows of the hot path whose arithmetic lives
in the reference checkout itself: step ordering, reward, waypoint progress,
termination, truncation, info, spaces and the SingleAgentWrapper output conventions.

How: `/root/reference/torchdriveenv/gym_env.py` is imported UNMODIFIED.  Its third-party imports that are
absent offline (gymnasium, invertedai, torchdrivesim) are satisfied by name-only stand-ins, and the one
function that glues the env to torchdrivesim (`build_simulator`, gym_env.py:179-300) is replaced by a
factory that returns an object with the SimulatorInterface-level call surface (step,
get_state, compute_offroad, compute_collision, compute_traffic_lights_violations, render_egocentric,
to, copy) whose state comes from this repo's CPU oracle.  Everything above that surface is the real
reference: `WaypointSuiteEnv.__init__/reset/set_start_pos/step/check_reach_target/get_reward/
is_terminated/is_truncated/get_info`, `GymEnv.step/get_obs`, `SingleAgentWrapper.step/transform_out`.

What it pins: given the same simulator states and infraction values, the reference's Python decides the
reward, flags, info and waypoint progress; the fixtures freeze those decisions, and tests compare the
oracle's (CPU) and the CUDA path's (GPU) own outputs for the same episode against them.  What it does
NOT pin: kinematics, collision, offroad, lights and the birdview themselves (torchdrivesim, absent).

Needs `/root/reference` (this container only; the fixtures travel).  Run from the repo root:
    python tests/golden/make_reference_golden.py
"""
import importlib
import json
import math
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REFERENCE = os.environ.get("TDE_REFERENCE", "/root/reference")

INFO_KEYS = ["offroad", "collision", "traffic_light_violation", "is_success", "reached_waypoint_num",
             "psi_smoothness", "psi_reward", "dist_reward", "speed_smoothness"]


# ----------------------------------------------------------------------------- name-only stand-ins
def _install_stand_ins():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class Anything:
        """Accepts any constructor arguments; attribute access yields another Anything."""
        def __init__(self, *a, **k):
            self.__dict__.update(k)

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return Anything()

        def __call__(self, *a, **k):
            return Anything()

    def cls(name):
        return type(name, (Anything,), {})

    # gymnasium: the three things gym_env.py touches
    class Env:
        metadata = {}

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name == "env":
                raise AttributeError(name)
            return getattr(self.env, name)

        def reset(self, **kwargs):
            return self.env.reset(**kwargs)

        def step(self, action):
            return self.env.step(action)

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.dtype = low, high, np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)

    registry = {}
    gym = mod("gymnasium", Env=Env, Wrapper=Wrapper, register=lambda id, entry_point=None, **k: registry.__setitem__(id, entry_point),
              registry=registry)
    gym.spaces = mod("gymnasium.spaces", Box=Box)

    mod("invertedai")
    mod("invertedai.common", AgentState=cls("AgentState"), Point=cls("Point"), AgentAttributes=cls("AgentAttributes"),
        RecurrentState=cls("RecurrentState"))
    mod("torchdrivesim")
    mod("torchdrivesim.behavior")
    mod("torchdrivesim.behavior.iai", IAIWrapper=cls("IAIWrapper"))
    mod("torchdrivesim.goals", WaypointGoal=cls("WaypointGoal"))
    mod("torchdrivesim.kinematic", KinematicBicycle=cls("KinematicBicycle"))
    mod("torchdrivesim.rendering", renderer_from_config=Anything())
    mod("torchdrivesim.rendering.base", RendererConfig=cls("RendererConfig"))
    mod("torchdrivesim.utils", Resolution=cls("Resolution"))
    mod("torchdrivesim.lanelet2", find_lanelet_directions=None)
    mod("torchdrivesim.map", find_map_config=None, traffic_controls_from_map_config=None)
    mod("torchdrivesim.traffic_lights", current_light_state_tensor_from_controller=None)
    mod("torchdrivesim.simulator", TorchDriveConfig=cls("TorchDriveConfig"), SimulatorInterface=cls("SimulatorInterface"),
        BirdviewRecordingWrapper=cls("BirdviewRecordingWrapper"), Simulator=cls("Simulator"),
        HomogeneousWrapper=cls("HomogeneousWrapper"), CollisionMetric=type("CollisionMetric", (), {"nograd": "nograd"}))


def import_reference():
    """Imports the reference's gym_env.py; the stand-ins only live in sys.modules while it is being imported
    (the module keeps the names it bound), so nothing else in the process ever sees them."""
    before = set(sys.modules)
    _install_stand_ins()
    sys.path.insert(0, REFERENCE)
    try:
        ref = importlib.import_module("torchdriveenv.gym_env")
    finally:
        sys.path.remove(REFERENCE)
        for name in set(sys.modules) - before:
            if name.split(".")[0] in ("gymnasium", "invertedai", "torchdrivesim"):
                del sys.modules[name]
    return ref


# ----------------------------------------------------------------------------- the simulator surface
class OracleSimulator:
    """SimulatorInterface-level surface over one env of the CPU oracle, B = A = 1 at the
    interface (NPCs hidden, as IAIWrapper hides them, gym_env.py:269-271)."""

    def __init__(self, orc):
        self.orc = orc
        self.outputs = None      # the oracle's own (reward, terminated, truncated, info row) of the last step

    def step(self, action):
        a = np.asarray(action.detach().cpu().numpy(), np.float32).reshape(-1)[:2]
        obs, rew, term, trunc, info = self.orc.step(a[None])
        self.outputs = (obs[0].copy(), float(rew[0]), bool(term[0]), bool(trunc[0]), info[0].copy())

    def get_state(self):
        return torch.from_numpy(self.orc.state[0:1, 0:1].copy())                      # 1 x 1 x 4

    def _infr(self, k):
        return torch.from_numpy(self.orc.infractions[0:1, 0, k].copy()).reshape(1, 1)  # 1 x 1

    def compute_collision(self): return self._infr(0)
    def compute_offroad(self): return self._infr(1)
    def compute_traffic_lights_violations(self): return self._infr(2)
    def compute_wrong_way(self): return self._infr(3)

    def render_egocentric(self):
        return torch.from_numpy(self.orc.render()[None].astype(np.float32))            # 1 x 1 x 3 x 64 x 64

    def to(self, device): return self
    def copy(self): return self


# ----------------------------------------------------------------------------- cases
def scenario_sets():
    from torchdriveenv_b200 import scenarios as S
    return {
        "three_way": (lambda: S.three_way(6), 9),
        "traffic_lights": (lambda: S.traffic_lights(12), 12),
        "roundabout": (lambda: S.roundabout(8), 8),
        "validation_mix": (lambda: S.validation_mix(8), 10),
    }


CASES = {
    # name: (scenario set, env-config overrides, steps, policy, numpy seed)
    "ref_three_way_default": ("three_way", dict(), 60, "pursuit", 11),
    "ref_three_way_no_termination": ("three_way", dict(terminated_at_infraction=False, distance_cutoff=0.25), 206, "pursuit", 12),
    "ref_traffic_lights_pursuit": ("traffic_lights", dict(terminated_at_infraction=False, distance_cutoff=0.25), 206, "pursuit", 13),
    "ref_roundabout_random": ("roundabout", dict(), 40, "random", 14),
    "ref_mix_custom_rewards": ("validation_mix", dict(terminated_at_infraction=False, waypoint_bonus=10.0, heading_penalty=5.0,
                                                      distance_bonus=0.5, distance_cutoff=0.1, max_environment_steps=50), 60, "pursuit", 15),
    "ref_mix_slow_start": ("validation_mix", dict(terminated_at_infraction=True, distance_cutoff=0.5), 80, "crawl", 16),
    "ref_three_way_swerve_offroad": ("three_way", dict(), 50, "swerve", 17),
    "ref_traffic_lights_swerve": ("traffic_lights", dict(distance_cutoff=0.25), 70, "swerve", 18),
    "ref_traffic_lights_red_run": ("traffic_lights", dict(terminated_at_infraction=True, distance_cutoff=0.25), 206, "pursuit", 19),
    "ref_three_way_oncoming": ("three_way", dict(), 120, "oncoming", 20),
}


def oracle_config(ref_cfg, A):
    """EnvConfig (the reference's dataclass instance) -> tde_config, through the product's own mapping."""
    from torchdriveenv_b200 import gym_env as G
    from torchdriveenv_b200._capi import default_config
    mine = G.EnvConfig(**{k: getattr(ref_cfg, k) for k in ("max_environment_steps", "waypoint_bonus", "heading_penalty", "distance_bonus",
                                                          "distance_cutoff", "terminated_at_infraction")})
    return default_config(num_envs=1, max_agents=A, **G.engine_config(mine, auto_reset=0))


def drive_episode(ref, cfg, data, ss, A, steps, policy, seed, check_build=None):
    """Runs the reference's SingleAgentWrapper(WaypointSuiteEnv(cfg, data)) for `steps` steps over an OracleSimulator on
    the scenario set `ss`; returns the recorded episode and the agreement report."""
    from oracle import oracle as O
    packed = ss.pack(A)
    made = {}

    # the glue to torchdrivesim, replaced (see the module docstring)
    locations = list(data.locations)      # find_map_config(f"carla_{location}") :312 -> the suite entry of that name
    ref.find_map_config = lambda map_name: types.SimpleNamespace(name=map_name, lanelet_map=locations.index(map_name[len("carla_"):]))
    ref.find_lanelet_directions = lambda lanelet_map, x, y: [float(ss.scenarios[lanelet_map].start_heading)]

    def build_simulator(cfg, map_cfg, device, ego_state, scenario=None, car_sequences=None, waypointseq=None):
        k = map_cfg.lanelet_map
        if check_build is not None:
            check_build(k, scenario, car_sequences, waypointseq, packed)
        orc = O.OracleEnvSet(oracle_config(cfg, A), packed)
        orc.set_env_scenario_range([k], [k + 1])
        orc.reset(seed=seed)
        start = np.asarray(ego_state, np.float32)
        orc.state[0, 0, :] = start
        made.update(orc=orc, scenario=k, start=start.copy(), waypointseq=waypointseq)
        return OracleSimulator(orc)
    ref.build_simulator = build_simulator

    env = ref.SingleAgentWrapper(ref.WaypointSuiteEnv(cfg=cfg, data=data))
    obs, _ = env.reset()
    assert obs.shape == (3, 64, 64) and obs.dtype == np.uint8
    assert env.action_space.shape == (2,) and env.observation_space.shape == (3, 64, 64)
    rng = np.random.default_rng(seed)
    out = dict(actions=[], reward=[], terminated=[], truncated=[], info=[], states=[], target_idx=[], orc_reward=[], orc_flags=[], orc_info=[])
    types_seen = {}
    for t in range(steps):
        sim = env.env.simulator
        x, y, psi, v = sim.orc.state[0, 0]
        tgt = env.env.current_target
        if policy == "random" or tgt is None:
            a = np.array([rng.uniform(-1, 1), rng.uniform(-0.3, 0.3)], np.float32)
        elif policy == "swerve":     # straight for a second, then a constant turn: leaves the road
            a = np.array([1.0, 0.0 if t < 10 else (0.12 if seed % 2 else -0.12)], np.float32)
        elif policy == "oncoming":   # drift into the oncoming lane (left-handed map: towards the lane beside the ego's)
            a = np.array([0.5, 0.0 if (t < 5 or t > 16) else (0.1 if t < 11 else -0.1)], np.float32)
        else:
            err = math.atan2(tgt[1] - y, tgt[0] - x) - psi
            err = (err + math.pi) % (2 * math.pi) - math.pi
            v_want = 0.6 if policy == "crawl" else 8.0
            a = np.array([np.clip(0.8 * (v_want - v), -1, 1), np.clip(0.35 * err + rng.normal(0, 0.02), -0.3, 0.3)], np.float32)
        obs, reward, terminated, truncated, info = env.step(a)
        o_obs, o_rew, o_term, o_trunc, o_info = sim.outputs
        assert np.array_equal(obs, o_obs)
        types_seen = dict(obs=[str(obs.dtype), list(obs.shape)], reward=type(reward).__name__, terminated=type(terminated).__name__,
                          truncated=type(truncated).__name__,
                          info={k: (type(val).__name__ if not torch.is_tensor(val) else f"tensor{list(val.shape)}") for k, val in info.items()})
        out["actions"].append(a)
        out["reward"].append(reward); out["terminated"].append(terminated); out["truncated"].append(truncated)
        out["info"].append([float(info[k]) for k in INFO_KEYS])
        out["states"].append(sim.orc.state[0, 0].copy())
        out["target_idx"].append(env.env.current_target_idx)
        out["orc_reward"].append(o_rew); out["orc_flags"].append([o_term, o_trunc]); out["orc_info"].append(o_info)
    res = dict(actions=np.asarray(out["actions"], np.float32), reward=np.asarray(out["reward"], np.float64),
               terminated=np.asarray(out["terminated"], np.uint8), truncated=np.asarray(out["truncated"], np.uint8),
               info=np.asarray(out["info"], np.float64), states=np.asarray(out["states"], np.float32),
               target_idx=np.asarray(out["target_idx"], np.int32), start_state=made["start"], scenario=np.int32(made["scenario"]),
               seed=np.int64(seed), max_agents=np.int32(A), info_keys=np.str_(json.dumps(INFO_KEYS)), types=np.str_(json.dumps(types_seen)))
    # the comparison itself (also what tests/test_reference_golden.py repeats from the frozen vectors)
    orc_info = np.asarray(out["orc_info"], np.float64)
    report = dict(reward_max_abs=float(np.max(np.abs(res["reward"] - np.asarray(out["orc_reward"])))),
                  flags_equal=bool(np.array_equal(np.stack([res["terminated"], res["truncated"]], 1), np.asarray(out["orc_flags"], np.uint8))),
                  info_max_abs=float(np.max(np.abs(res["info"] - orc_info[:, :len(INFO_KEYS)]))),
                  waypoints_reached=int(res["info"][-1, 4]), terminated_steps=int(res["terminated"].sum()),
                  offroad_steps=int((res["info"][:, 0] > 0).sum()), collision_steps=int((res["info"][:, 1] > 0).sum()),
                  red_light_steps=int((res["info"][:, 2] > 0).sum()),
                  truncated_steps=int(res["truncated"].sum()))
    return res, report


def run_case(ref, name):
    set_name, overrides, steps, policy, seed = CASES[name]
    builder, A = scenario_sets()[set_name]
    ss = builder()
    cfg = ref.EnvConfig(seed=seed, **overrides)
    data = ref.WaypointSuite(locations=[f"Town{k:02d}" for k in range(len(ss.scenarios))],
                             waypoint_suite=[np.asarray(s.waypoints, np.float64).tolist() for s in ss.scenarios],
                             car_sequence_suite=[None] * len(ss.scenarios), scenarios=[None] * len(ss.scenarios))
    res, report = drive_episode(ref, cfg, data, ss, A, steps, policy, seed)
    res.update(scenario_set=np.str_(set_name), env_config=np.str_(json.dumps(overrides)))
    return res, report


# ----------------------------------------------------------------------------- the reference's own validation suite
VALIDATION_YML = os.path.join(REFERENCE, "torchdriveenv", "data", "validation_cases.yml")
SUITE_CASES = {   # name: (entry of validation_cases.yml, env-config overrides, steps, policy, seed)
    "refsuite_case0_three_way": (0, dict(terminated_at_infraction=False), 150, "pursuit", 31),
    "refsuite_case1_parked_car": (1, dict(terminated_at_infraction=False), 150, "pursuit", 32),
    "refsuite_case2_chicken": (2, dict(), 100, "pursuit", 33),
    "refsuite_case3_roundabout": (3, dict(terminated_at_infraction=False, distance_cutoff=0.25), 206, "pursuit", 34),
    "refsuite_case4_traffic_lights": (4, dict(terminated_at_infraction=False, distance_cutoff=0.25), 206, "pursuit", 35),
}


def import_reference_env_utils():
    """The reference's torchdriveenv/env_utils.py, unmodified; OmegaConf (absent) is stood in for by PyYAML."""
    import yaml
    before = set(sys.modules)
    _install_stand_ins()
    om = types.ModuleType("omegaconf")
    om.OmegaConf = type("OmegaConf", (), {"load": staticmethod(lambda path: yaml.load(open(path), Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))),
                                          "to_object": staticmethod(lambda x: x)})
    sys.modules["omegaconf"] = om
    sys.path.insert(0, REFERENCE)
    try:
        mod = importlib.import_module("torchdriveenv.env_utils")
    finally:
        sys.path.remove(REFERENCE)
        for name in set(sys.modules) - before:
            if name.split(".")[0] in ("gymnasium", "invertedai", "torchdrivesim", "omegaconf"):
                del sys.modules[name]
    return mod


def suite_entry(data, k):
    """Entry k of a WaypointSuite as plain arrays (what the fixture stores and the tests rebuild the scenario from)."""
    sc = data.scenarios[k]
    seqs = data.car_sequence_suite[k] or {}
    keys = sorted(int(q) for q in seqs)
    return dict(suite_location=np.str_(data.locations[k]), suite_waypoints=np.asarray(data.waypoint_suite[k], np.float64),
                suite_agent_states=np.asarray(sc.agent_states if sc is not None else [], np.float64).reshape(-1, 4),
                suite_agent_attributes=np.asarray(sc.agent_attributes if sc is not None else [], np.float64).reshape(-1, 3),
                suite_car_seq_keys=np.asarray(keys, np.int32),
                suite_car_seqs=(np.asarray([seqs[q] if q in seqs else seqs[str(q)] for q in keys], np.float64).reshape(len(keys), -1, 4)
                                if keys else np.zeros((0, 0, 4), np.float64)))


def suite_from_entry(d, G):
    """The inverse of suite_entry: a one-entry WaypointSuite of the product's dataclasses."""
    sc = None
    if len(d["suite_agent_states"]):
        sc = G.Scenario(agent_states=d["suite_agent_states"].tolist(), agent_attributes=d["suite_agent_attributes"].tolist(),
                        recurrent_states=[[0.0]] * len(d["suite_agent_states"]))
    seqs = {int(q): d["suite_car_seqs"][i].tolist() for i, q in enumerate(d["suite_car_seq_keys"])}
    return G.WaypointSuite(locations=[str(d["suite_location"])], waypoint_suite=[d["suite_waypoints"].tolist()],
                           car_sequence_suite=[seqs], scenarios=[sc])


def run_suite_case(ref, ref_utils, name):
    """One entry of the reference's validation_cases.yml, loaded by the REFERENCE'S loader, driven through the reference's
    env; the scenario tables come from the product's loader + scenario_set_from_suite on the same file, and the
    arguments the reference hands to build_simulator are checked against those tables."""
    from torchdriveenv_b200 import env_utils as U, gym_env as G
    k, overrides, steps, policy, seed = SUITE_CASES[name]
    data_ref = ref_utils.load_waypoint_suite_data(VALIDATION_YML)
    data_mine = U.load_waypoint_suite_data(VALIDATION_YML)
    for field in ("locations", "waypoint_suite", "car_sequence_suite"):
        assert getattr(data_ref, field) == getattr(data_mine, field), field
    for a, b in zip(data_ref.scenarios, data_mine.scenarios):
        assert (a is None) == (b is None) and (a is None or (a.agent_states == b.agent_states and a.agent_attributes == b.agent_attributes))
    entry = suite_entry(data_mine, k)
    one_mine = suite_from_entry(entry, G)
    one_ref = ref.WaypointSuite(locations=[data_ref.locations[k]], waypoint_suite=[data_ref.waypoint_suite[k]],
                                car_sequence_suite=[data_ref.car_sequence_suite[k]], scenarios=[data_ref.scenarios[k]])
    cfg = ref.EnvConfig(seed=seed, **overrides)
    ss = G.scenario_set_from_suite(G.EnvConfig(**overrides), one_mine, n_background=0, seed=0)
    A = ss.max_agents()

    def check_build(idx, scenario, car_sequences, waypointseq, packed):
        # what the reference passes down (gym_env.py:339-345) against the tables the product built from the same file
        assert idx == 0
        assert np.allclose(packed["waypoints"], np.asarray(waypointseq, np.float32))
        n_pre = 0 if scenario is None else len(scenario.agent_states)
        if n_pre:
            assert np.allclose(packed["agent_init"][0, 1:1 + n_pre], np.asarray(scenario.agent_states, np.float32))
            assert np.allclose(packed["agent_attr"][0, 1:1 + n_pre], np.asarray(scenario.agent_attributes, np.float32))
        assert int(packed["scen_num_agents"][0]) == 1 + n_pre
        for key, seq in (car_sequences or {}).items():      # dict key = agent slot (gym_env.py:279)
            seq = np.asarray(seq, np.float32)
            rs = packed["replay_states"].reshape(-1, A, 4)
            assert np.allclose(rs[: len(seq), int(key)], seq) and packed["replay_mask"][: len(seq), int(key)].all()
    res, report = drive_episode(ref, cfg, one_ref, ss, A, steps, policy, seed, check_build)
    res.update(entry)
    res.update(env_config=np.str_(json.dumps(overrides)), suite_index=np.int32(k))
    return res, report


# ----------------------------------------------------------------------------- the reference's episode metrics
def import_reference_eval_callback():
    """examples/rl_training.py, unmodified, for its EvalNTimestepsCallback._calc_metrics (:39-67): the per-episode
    aggregates the episode-statistics vector of the C ABI (TDE_STAT_*) stands for.  stable-baselines3 and wandb
    (absent) are name-only stand-ins; the callback's metric code does not touch them."""
    import yaml
    before = set(sys.modules)
    _install_stand_ins()

    def mod(name, **attrs):
        m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m
        return m
    blank = lambda n: type(n, (), {"__init__": lambda self, *a, **k: None})
    mod("omegaconf", OmegaConf=type("OmegaConf", (), {"load": staticmethod(lambda path: yaml.load(open(path), Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))),
                                                      "to_object": staticmethod(lambda x: x)}))
    mod("wandb"); mod("wandb.integration"); mod("wandb.integration.sb3", WandbCallback=blank("WandbCallback"))
    mod("stable_baselines3", SAC=blank("SAC"), PPO=blank("PPO"), A2C=blank("A2C"), TD3=blank("TD3"))
    mod("stable_baselines3.common")
    mod("stable_baselines3.common.monitor", Monitor=blank("Monitor"))
    mod("stable_baselines3.common.vec_env", VecVideoRecorder=blank("VecVideoRecorder"), VecFrameStack=blank("VecFrameStack"), SubprocVecEnv=blank("SubprocVecEnv"))
    mod("stable_baselines3.common.callbacks", BaseCallback=blank("BaseCallback"))
    mod("stable_baselines3.common.evaluation", evaluate_policy=None)
    ex = os.path.join(REFERENCE, "examples")
    sys.path.insert(0, REFERENCE); sys.path.insert(0, ex)
    try:
        rl = importlib.import_module("rl_training")
    finally:
        sys.path.remove(ex); sys.path.remove(REFERENCE)
        for name in set(sys.modules) - before:
            if name.split(".")[0] in ("gymnasium", "invertedai", "torchdrivesim", "omegaconf", "wandb", "stable_baselines3", "common", "rl_training"):
                del sys.modules[name]
    return rl


def reference_episode_metrics(rl, infos_per_episode):
    """Feeds per-step info dicts (one list per episode) through the reference's EvalNTimestepsCallback._calc_metrics,
    the way _evaluate does (:83-95), and returns its counters."""
    cb = rl.EvalNTimestepsCallback(eval_env=None, n_steps=1, eval_n_episodes=len(infos_per_episode))
    cb.episode_num = cb.offroad_num = cb.collision_num = cb.traffic_light_violation_num = cb.success_num = 0
    cb.reached_waypoint_nums, cb.psi_smoothness, cb.speed_smoothness = [], [], []
    for infos in infos_per_episode:
        cb.psi_smoothness_for_single_episode, cb.speed_smoothness_for_single_episode = [], []
        for info in infos:
            cb._calc_metrics({"info": info}, {})
    return dict(episodes=cb.episode_num, offroad=cb.offroad_num, collision=cb.collision_num,
                traffic_light_violation=cb.traffic_light_violation_num, success=cb.success_num,
                reached_waypoints=float(sum(cb.reached_waypoint_nums)))


def run_metrics_case(rl, E=48, steps=260, seed=41):
    """E oracle envs with auto-reset, random actions: every finished episode's per-step infos go through the reference's
    metric code; its counters must equal the oracle's episode-statistics vector."""
    from oracle import oracle as O
    from torchdriveenv_b200 import scenarios as S
    from torchdriveenv_b200._capi import INFO_COLUMNS as IC, STAT_NAMES, default_config
    ss = S.validation_mix(8)
    A = ss.max_agents()
    orc = O.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1, max_environment_steps=60), ss.pack(A))
    orc.reset(seed=seed)
    rng = np.random.default_rng(seed)
    open_eps = [[] for _ in range(E)]
    done_eps = []
    acts = []
    for t in range(steps):
        a = np.stack([rng.uniform(-0.2, 1, E), rng.uniform(-0.1, 0.1, E)], 1).astype(np.float32)
        acts.append(a)
        _, rew, term, trunc, info = orc.step(a)
        for e in range(E):
            row = info[e]
            open_eps[e].append(dict(offroad=row[IC["offroad"]], collision=row[IC["collision"]], traffic_light_violation=row[IC["traffic_light_violation"]],
                                    is_success=bool(row[IC["is_success"]] != 0), reached_waypoint_num=int(row[IC["reached_waypoint_num"]]),
                                    psi_smoothness=float(row[IC["psi_smoothness"]]), speed_smoothness=float(row[IC["speed_smoothness"]])))
            if term[e] or trunc[e]:
                done_eps.append(open_eps[e]); open_eps[e] = []
    want = reference_episode_metrics(rl, done_eps)
    stats = {n: float(orc.stats[i]) for i, n in enumerate(STAT_NAMES)}
    return dict(actions=np.asarray(acts, np.float32), seed=np.int64(seed), num_envs=np.int32(E), max_agents=np.int32(A),
                ref_metrics=np.str_(json.dumps(want))), want, stats


# ----------------------------------------------------------------------------- scenario-builder JSON (row f2)
LABELED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "labeled_json")


def write_labeled_inputs():
    """Three small scenario-builder files with the schema the reference reads (env_utils.py:31-105): a plain route, a
    route with a parked car (max_speed 0 -> 200-step replay) and a moving one (5 states -> replay), and a route with a
    single-state agent (random initial speed 5..10: the only random draw, so the file order does not matter)."""
    os.makedirs(LABELED_DIR, exist_ok=True)
    st = lambda x, y, o=0.0: dict(center=dict(x=x, y=y), orientation=o, speed=0.0)
    route = lambda pts: {"0": dict(states=[st(x, y) for x, y in pts])}
    attrs = lambda l, w, r, **k: dict(length=l, width=w, rear_axis_offset=r, **k)
    files = {
        "carla_Town01_plain.json": dict(individual_suggestions=route([(0, 0), (12, 0), (24, 3), (36, 9)]), predetermined_agents=None),
        "carla_Town03_replay.json": dict(individual_suggestions=route([(5, 5), (5, 20), (5, 35)]),
                                         predetermined_agents={"1": dict(states={"0": st(5.0, 30.0, 1.57)}, static_attributes=attrs(4.9, 2.0, 1.4, max_speed=0)),
                                                               "2": dict(states={str(i): st(9.0, 10.0 + 2.0 * i, 1.57) for i in range(5)},
                                                                         static_attributes=attrs(5.2, 2.1, 1.5))}),
        "carla_Town07_single.json": dict(individual_suggestions=route([(-3, 1), (-15, 1), (-27, 4)]),
                                         predetermined_agents={"1": dict(states={"0": st(-20.0, 4.5, 3.14)}, static_attributes=attrs(4.5, 1.9, 1.3))}),
    }
    for name, content in files.items():
        with open(os.path.join(LABELED_DIR, name), "w") as f:
            json.dump(content, f, indent=1)


def suite_as_plain(suite):
    """WaypointSuite (either package's dataclass) -> JSON-able entries sorted by location."""
    rows = []
    for k, loc in enumerate(suite.locations):
        sc = suite.scenarios[k]
        seqs = suite.car_sequence_suite[k]
        rows.append(dict(location=loc, waypoints=suite.waypoint_suite[k],
                         agent_states=None if sc is None else sc.agent_states, agent_attributes=None if sc is None else sc.agent_attributes,
                         n_recurrent=None if sc is None else [len(r) for r in sc.recurrent_states],
                         car_sequences=None if seqs is None else {str(q): v for q, v in sorted(seqs.items())}))
    return sorted(rows, key=lambda r: r["location"])


def run_labeled_case(ref_utils):
    import random
    write_labeled_inputs()
    random.seed(7)
    return suite_as_plain(ref_utils.load_labeled_data(LABELED_DIR))


if __name__ == "__main__":
    ref = import_reference()
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        res, report = run_case(ref, name)
        np.savez_compressed(os.path.join(here, name + ".npz"), **res)
        print(name, report)
    rl = import_reference_eval_callback()
    res, want, stats = run_metrics_case(rl)
    print("refmetrics_validation_mix", want, {k: stats[k] for k in want})
    assert all(float(want[k]) == stats[k] for k in want)
    np.savez_compressed(os.path.join(here, "refmetrics_validation_mix.npz"), **res)
    ref_utils = import_reference_env_utils()
    # EnvConfig as the reference constructs it from the `env:` section of its shipped training configs (env_utils.py:10-12)
    import glob
    import yaml
    env_cfgs = {}
    for path in sorted(glob.glob(os.path.join(REFERENCE, "examples", "env_configs", "*", "*.yml"))):
        raw = yaml.safe_load(open(path))["env"]
        c = ref_utils.construct_env_config(raw)
        env_cfgs["/".join(path.split(os.sep)[-2:])] = dict(raw=raw, fields={k: getattr(c, k) for k in c.__dataclass_fields__ if k != "simulator"})
    with open(os.path.join(here, "ref_env_configs.json"), "w") as f:
        json.dump(env_cfgs, f, indent=1)
    print("ref_env_configs", len(env_cfgs))
    labeled = run_labeled_case(ref_utils)
    with open(os.path.join(here, "ref_labeled_suite.json"), "w") as f:
        json.dump(labeled, f, indent=1)
    print("ref_labeled_suite", [(r["location"], None if r["agent_states"] is None else len(r["agent_states"])) for r in labeled])
    for name in SUITE_CASES:
        res, report = run_suite_case(ref, ref_utils, name)
        np.savez_compressed(os.path.join(here, name + ".npz"), **res)
        print(name, report)
