"""gym.Env-level API of the reference, served by the CUDA engine.

Mirrors torchdriveenv/gym_env.py of the reference name for name: ``EnvConfig`` :34-54, ``Scenario``
:56-60, ``WaypointSuite`` :63-68, ``GymEnv`` :71-177, ``build_simulator`` :179-300,
``WaypointSuiteEnv`` :303-437, ``SingleAgentWrapper`` :440-487.  What the reference computes per step
in Python on a B=1 torchdrivesim simulator (reward :396-411, termination :413-417, truncation
:134-135, info :419-437, waypoint progress :378-394) is computed inside the fused CUDA step kernel;
this module only marshals tensors.  ``TorchDriveVecEnv`` is the batched sibling (SB3 VecEnv shape,
examples/rl_training.py:159-160) that steps thousands of envs in lockstep.

gymnasium / stable-baselines3 are optional: when absent the spaces are small stand-ins.
"""
from __future__ import annotations

import logging
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from ._capi import INFO_COLUMNS, TDE_OBS_H, TDE_OBS_W
from .engine import Engine
from .scenarios import (MapData, ScenarioData, ScenarioSet, build_polyline_map, place_npcs)
from .helpers import save_video
from .simulator import BatchedSimulator, BirdviewRecordingWrapper

logger = logging.getLogger(__name__)

try:  # pragma: no cover - gymnasium is not installed in the build image
    import gymnasium as gym
    _Env, _Wrapper, _Box = gym.Env, gym.Wrapper, gym.spaces.Box
except Exception:  # minimal stand-ins with the attributes the reference uses
    gym = None

    class _Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, self.dtype), self.shape).copy()

        def sample(self):
            if np.issubdtype(self.dtype, np.integer):
                return np.random.randint(self.low, self.high.astype(np.int64) + 1).astype(self.dtype)
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class _Env:
        metadata: Dict = {}

    class _Wrapper(_Env):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            return getattr(self.env, name)

        def reset(self, **kwargs):
            return self.env.reset(**kwargs)

        def step(self, action):
            return self.env.step(action)


@dataclass
class SimulatorConfig:
    """The TorchDriveConfig / RendererConfig values the reference env relies on (gym_env.py:46-49)."""
    left_handed_coordinates: bool = True
    highlight_ego_vehicle: bool = True
    offroad_threshold: float = 0.5
    fov: float = 35.0
    tl_rear_factor: float = 0.1


@dataclass
class EnvConfig:
    ego_only: bool = False
    max_environment_steps: int = 200
    frame_stack: int = 3
    waypoint_bonus: float = 100.
    heading_penalty: float = 25.
    distance_bonus: float = 1.
    distance_cutoff: float = 0.5
    use_background_traffic: bool = True
    terminated_at_infraction: bool = True
    seed: Optional[int] = None
    simulator: SimulatorConfig = field(default_factory=SimulatorConfig)
    render_mode: Optional[str] = "rgb_array"
    video_filename: Optional[str] = "rendered_video.mp4"
    video_res: Optional[int] = 1024
    video_fov: Optional[float] = 500
    device: Optional[str] = None


@dataclass
class Scenario:
    agent_states: List[List[float]] = None
    agent_attributes: List[List[float]] = None
    recurrent_states: List[List[float]] = None


@dataclass
class WaypointSuite:
    locations: List[str] = None
    waypoint_suite: List[List[List[float]]] = None
    car_sequence_suite: List[Optional[Dict[int, List[List[float]]]]] = None
    scenarios: List[Optional[Scenario]] = None


def engine_config(cfg: EnvConfig, **extra) -> Dict:
    sim = cfg.simulator
    out = dict(max_environment_steps=int(cfg.max_environment_steps), waypoint_bonus=float(cfg.waypoint_bonus),
               heading_penalty=float(cfg.heading_penalty), distance_bonus=float(cfg.distance_bonus),
               distance_cutoff=float(cfg.distance_cutoff), terminated_at_infraction=int(bool(cfg.terminated_at_infraction)),
               left_handed_coordinates=int(bool(sim.left_handed_coordinates)), offroad_threshold=float(sim.offroad_threshold),
               fov=float(sim.fov), tl_rear_factor=float(sim.tl_rear_factor))
    out.update(extra)
    return out


def scenario_set_from_suite(cfg: EnvConfig, data: WaypointSuite, n_background: int = 0, seed: int = 0,
                            map_builder=None, background_traffic=None) -> ScenarioSet:
    """Scenario tables from a WaypointSuite, following build_simulator (gym_env.py:179-300): waypoints
    :252-257, predetermined agents :222-228, replay tensors from car sequences :275-283.  The CARLA map
    assets (find_map_config :312) are not available offline, so each entry gets a synthetic lane mesh
    extruded from its own waypoint polyline; the Inverted AI background agents (:236-238) are replaced
    by ``n_background`` constant-speed replay NPCs.  ``background_traffic`` (a dict from
    env_utils.load_background_traffic, or a list with one per suite entry) adds the file's agents that
    are farther than 100 m from the start (:229-233) as constant-velocity NPCs."""
    maps: List[MapData] = []
    scen: List[ScenarioData] = []
    rng = np.random.default_rng(seed)
    builder = map_builder or (lambda poly, name: build_polyline_map(poly, name, with_lights=True))
    for k, poly in enumerate(data.waypoint_suite):
        poly = np.asarray(poly, np.float64)
        name = f"{(data.locations[k] if data.locations else 'map')}_{k}"
        maps.append(builder(poly, name))
        d = poly[1] - poly[0]
        heading = float(np.arctan2(d[1], d[0]))
        states = [[poly[0][0], poly[0][1], heading, 0.0]]
        attrs = [[5.0, 2.0, 0.9]]
        sc = data.scenarios[k] if data.scenarios is not None else None
        if not cfg.ego_only and sc is not None and sc.agent_states is not None:
            states += [list(map(float, s)) for s in sc.agent_states]
            attrs += [list(map(float, a)) for a in sc.agent_attributes]
        n_pre = len(states)
        seqs = data.car_sequence_suite[k] if (data.car_sequence_suite is not None and not cfg.ego_only) else None
        bg_states = bg_attrs = bg_rep = bg_mask = None
        if not cfg.ego_only and n_background > 0:
            bg_states, bg_attrs, bg_rep, bg_mask = place_npcs(poly, n_background, rng)
            states += bg_states.tolist(); attrs += bg_attrs.tolist()
        if not cfg.ego_only and cfg.use_background_traffic and background_traffic is not None:
            from .env_utils import background_agents_for_start
            bt = background_traffic[k] if isinstance(background_traffic, (list, tuple)) else background_traffic
            if bt is not None:
                bst, bat = background_agents_for_start(bt, poly[0], max_agents=max(0, 64 - len(states)))
                states += bst.tolist(); attrs += bat.tolist()
        n = len(states)
        T = 0
        if seqs:
            T = max(len(v) for v in seqs.values())
        if bg_rep is not None:
            T = max(T, bg_rep.shape[0])
        rs = rm = None
        if T > 0:
            rs = np.zeros((T, n, 4), np.float32); rm = np.zeros((T, n), np.uint8)
            if seqs:
                for idx, seq in seqs.items():  # dict key = agent slot (gym_env.py:279)
                    idx = int(idx)
                    if 0 < idx < n and len(seq) > 0:
                        arr = np.asarray(seq, np.float32).reshape(-1, 4)
                        rs[: arr.shape[0], idx] = arr; rm[: arr.shape[0], idx] = 1
            if bg_rep is not None:
                rs[: bg_rep.shape[0], n_pre:] = bg_rep; rm[: bg_rep.shape[0], n_pre:] = bg_mask
        scen.append(ScenarioData(map_index=k, waypoints=poly.astype(np.float32), start_heading=heading,
                                 agent_init=np.asarray(states, np.float32), agent_attr=np.asarray(attrs, np.float32),
                                 replay_states=rs, replay_mask=rm, name=name))
    return ScenarioSet(maps, scen)


def _info_from_row(row: np.ndarray, as_tensor_device=None) -> Dict:
    """get_info (gym_env.py:419-437): infraction values as B x A tensors, the rest as Python scalars."""
    def t(v):
        x = torch.tensor([[float(v)]])
        return x.to(as_tensor_device) if as_tensor_device is not None else x
    c = INFO_COLUMNS
    return dict(
        offroad=t(row[c["offroad"]]), collision=t(row[c["collision"]]),
        traffic_light_violation=t(row[c["traffic_light_violation"]]),
        is_success=bool(row[c["is_success"]] != 0), reached_waypoint_num=int(row[c["reached_waypoint_num"]]),
        psi_smoothness=float(row[c["psi_smoothness"]]), psi_reward=float(row[c["psi_reward"]]),
        dist_reward=float(row[c["dist_reward"]]), speed_smoothness=float(row[c["speed_smoothness"]]),
        wrong_way=float(row[c["wrong_way"]]),
    )


class GymEnv(_Env):
    metadata = {"render_modes": ["video", "rgb_array"], "render_fps": 10}

    def __init__(self, cfg: EnvConfig, simulator):
        if cfg.render_mode is not None and cfg.render_mode not in self.metadata["render_modes"]:
            raise NotImplementedError
        self.render_mode = cfg.render_mode
        action_range = np.ndarray(shape=(2, 2), dtype=np.float32)
        action_range[:, 0] = (-1.0, 1.0)   # acceleration (gym_env.py:83)
        action_range[:, 1] = (-0.3, 0.3)   # steering (gym_env.py:84)
        self.max_environment_steps = cfg.max_environment_steps
        self.environment_steps = 0
        self.action_space = _Box(low=action_range[0], high=action_range[1], dtype=np.float32)
        self.observation_space = _Box(low=0, high=255, shape=(3, TDE_OBS_H, TDE_OBS_W), dtype=np.uint8)
        self.reward_range = (-float('inf'), float('inf'))
        self.collision_threshold = 0.0
        self.offroad_threshold = 0.0
        self.config = cfg
        self.simulator = simulator
        self.current_action = None
        self.last_birdview = None

    # -- the generic env of the reference (gym_env.py:107-150,159-170): whatever simulator it is given, driven through the
    #    SimulatorInterface-level surface; WaypointSuiteEnv overrides the per-step part with the fused CUDA step
    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        self.simulator = self.start_sim.copy()          # :110 (start_sim is set by the code that builds the env)
        self.environment_steps = 0
        self.last_birdview = None
        return self.get_obs(), {}

    def step(self, action):
        self.environment_steps += 1
        self.simulator.step(action)
        self.last_action = self.current_action if self.current_action is not None else action
        self.current_action = action
        return self.get_obs(), self.get_reward(), self.is_terminated(), self.is_truncated(), self.get_info()

    def get_obs(self):
        return self.simulator.render_egocentric().cpu().numpy().astype(np.uint8)

    def get_reward(self):
        x = self.simulator.get_state()[..., 0]
        return np.zeros(tuple(x.shape))

    def is_done(self):
        return self.is_truncated() or self.is_terminated()

    def is_truncated(self):
        return self.environment_steps >= self.max_environment_steps

    def is_terminated(self):
        return False

    def get_info(self):
        self.info = dict(
            offroad=self.simulator.compute_offroad(),
            collision=self.simulator.compute_collision(),
            traffic_light_violation=self.simulator.compute_traffic_lights_violations(),
            is_success=(self.environment_steps >= self.max_environment_steps),
        )
        return self.info

    def mock_step(self):
        obs = np.zeros((1, 3, TDE_OBS_H, TDE_OBS_W))
        info = dict(offroad=torch.Tensor([[0]]), collision=torch.Tensor([[0]]), traffic_light_violation=torch.Tensor([[0]]),
                    is_success=False)
        return obs, 0, False, True, info

    def seed(self, seed=None):
        pass

    def render(self):
        if self.render_mode == 'rgb_array':
            birdview = self.simulator.render_egocentric().cpu().numpy()
            return np.transpose(birdview.squeeze(), axes=(1, 2, 0))
        raise NotImplementedError

    def close(self):
        sim = getattr(self, "simulator", None)
        if isinstance(sim, BirdviewRecordingWrapper):    # gym_env.py:170-177
            bvs = sim.get_birdviews()
            if len(bvs) > 1:
                save_video(bvs, self.config.video_filename)
        if sim is not None and hasattr(sim, "engine"):
            sim.engine.close()


def build_simulator(cfg: EnvConfig, scenarios: ScenarioSet, device, num_envs: int = 1, seed: int = 0,
                    **extra) -> BatchedSimulator:
    """Counterpart of build_simulator (gym_env.py:179-300): the scenario tables play the role of the
    map config + agent tensors, the returned object exposes the SimulatorInterface call surface."""
    simulator = BatchedSimulator(scenarios, num_envs=num_envs, device=device, seed=seed, **engine_config(cfg, **extra))
    if cfg.render_mode == "video":   # gym_env.py:295-297
        simulator = BirdviewRecordingWrapper(simulator, res=(int(cfg.video_res), int(cfg.video_res)), fov=float(cfg.video_fov), to_cpu=True)
    return simulator


class WaypointSuiteEnv(GymEnv):
    """Single environment with the reference's API (B = A = 1 at the interface)."""

    def __init__(self, cfg: EnvConfig, data, n_background: int = 0):
        self.config = cfg
        if cfg.device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("torchdriveenv_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")
            self.torch_device = torch.device('cuda')
        else:
            self.torch_device = torch.device(cfg.device)
        self._seed = int(cfg.seed) if cfg.seed is not None else int(np.random.randint(0, 2**31 - 1))
        logger.info(f"seed: {self._seed}")
        self.scenario_set = data if isinstance(data, ScenarioSet) else scenario_set_from_suite(cfg, data, n_background, self._seed)
        super().__init__(cfg=cfg, simulator=None)
        self.simulator = build_simulator(cfg, self.scenario_set, str(self.torch_device), num_envs=1, seed=self._seed)
        self.engine: Engine = self.simulator.engine
        self.reached_waypoint_num = 0
        self.last_obs = self.last_reward = self.last_info = None

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        if seed is not None:
            # gymnasium's seeding contract: reset(seed=s) starts a reproducible stream - the same s gives the same
            # scenario, start pose and light phase.  The draws are keyed by (seed, env, episode counter): restart the counter.
            self._seed = int(seed)
            v = self.engine.get_env_vars()
            v[:, 5] = 0
            self.engine.set_env_vars(v)
        self.engine.reset(seed=self._seed)   # without a seed the engine's episode counter varies the draw per episode
        self.environment_steps = 0
        self.reached_waypoint_num = 0
        self.last_obs = self.last_reward = self.last_info = None
        if isinstance(self.simulator, BirdviewRecordingWrapper):   # the reference builds a fresh recorder per episode
            self.simulator.birdviews = []
            self.simulator._record()
        v = self.engine.get_env_vars()[0].cpu().numpy()
        self.current_waypoint_suite_idx = int(v[0])
        self.current_target_idx = int(v[2])
        return self.get_obs(), {}

    def get_obs(self):
        return self.simulator.render_egocentric().cpu().numpy().astype(np.uint8)

    # get_info's infraction entries are B x A tensors (gym_env.py:419-437).  They are made on the host by default: one env
    # behind this API is a latency path, and every tensor moved to the GPU and back (SingleAgentWrapper.transform_out) is a
    # synchronisation.  True puts them on the simulator's device, as the reference's simulator returns them.
    info_tensors_on_device = False

    def step(self, action):
        a = np.asarray(action.cpu() if torch.is_tensor(action) else action, dtype=np.float32).reshape(-1)[:2]
        # tde_step_host: actions up, one step, observation / reward / flags / info row down, ONE synchronisation
        obs, rew, term, trunc, info = self.engine.step_host(a.reshape(1, 2))
        getattr(self.simulator, "simulator", self.simulator)._infractions_valid = True   # on the wrapped simulator in video mode
        if isinstance(self.simulator, BirdviewRecordingWrapper):
            self.simulator._record()
        self.environment_steps += 1
        obs_np = obs.reshape(1, 1, 3, TDE_OBS_H, TDE_OBS_W).copy()       # the engine's pinned buffers are reused by the next call
        row = info[0].copy()
        reward = float(rew[0])
        terminated, truncated = bool(term[0]), bool(trunc[0])
        info_d = _info_from_row(row, self.torch_device if self.info_tensors_on_device else None)
        self.reached_waypoint_num = info_d["reached_waypoint_num"]
        self.last_obs, self.last_reward, self.last_info = obs_np, reward, info_d
        return obs_np, reward, terminated, truncated, info_d

    def is_terminated(self):
        if not self.config.terminated_at_infraction:
            return False
        s = self.simulator
        return bool(((s.compute_offroad() > 0) | (s.compute_collision() > 0) | (s.compute_traffic_lights_violations() > 0)).item())


class SingleAgentWrapper(_Wrapper):
    """Removes batch and agent dimensions from the environment interface (gym_env.py:440-487)."""

    def __init__(self, env):
        super().__init__(env)

    def reset(self, **kwargs):
        obs, _ = self.env.reset(**kwargs)
        return self.transform_out(obs), _

    def step(self, action):
        action = torch.Tensor(np.asarray(action, dtype=np.float32)).unsqueeze(0).unsqueeze(0)
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.transform_out(obs), self.transform_out(reward), self.transform_out(terminated), truncated, self.transform_out(info)

    def transform_out(self, x):
        if torch.is_tensor(x):
            t = x.squeeze(0).squeeze(0).cpu()
        elif isinstance(x, dict):
            t = {k: self.transform_out(v) for (k, v) in x.items()}
        elif isinstance(x, np.ndarray):
            t = self.transform_out(torch.tensor(x)).cpu().numpy()
        else:
            t = x
        return t

    def render(self, *args, **kwargs):
        return self.env.render(*args, **kwargs)

    def close(self):
        self.env.close()


try:  # pragma: no cover - stable-baselines3 is not installed in the build image
    from stable_baselines3.common.vec_env import VecEnv as _SB3VecEnv
except Exception:
    _SB3VecEnv = None


class VecInfos:
    """SB3's ``infos`` of one vectorised step: a sequence of E dicts, built on demand from the step's arrays.

    ``infos[i]`` is what ``SingleAgentWrapper.step`` returns for one env (get_info gym_env.py:419-437 through
    transform_out :463-472: ``offroad / collision / traffic_light_violation`` as scalars, ``is_success``,
    ``reached_waypoint_num``, ``psi_smoothness``, ``psi_reward``, ``dist_reward``, ``speed_smoothness``) - the dict that
    ``EvalNTimestepsCallback._calc_metrics`` (examples/rl_training.py:39-67) reads - plus what SB3's own wrappers add
    for an env that finished in this step: ``episode = {r, l, t}`` (``Monitor``, rl_training.py:123),
    ``TimeLimit.truncated`` and ``terminal_observation`` (``SubprocVecEnv``, :159).  ``infos.columns[name]`` is the
    whole column as one array (the fast path: no per-env Python objects)."""

    def __init__(self, info, terminated, truncated, terminal_observation=None, elapsed: float = 0.0):
        self._info, self._term, self._trunc, self._tobs, self._elapsed = info, terminated, truncated, terminal_observation, float(elapsed)
        self._host = None

    @property
    def columns(self) -> Dict:
        cols = {k: self._info[:, i] for k, i in INFO_COLUMNS.items()}
        cols["terminated"], cols["truncated"] = self._term, self._trunc
        if self._tobs is not None:
            cols["terminal_observation"] = self._tobs
        return cols

    def _arrays(self):
        if self._host is None:   # one device-to-host copy for the whole step, on first per-env access
            to_np = lambda t: t.cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
            self._host = (to_np(self._info), to_np(self._term).astype(bool), to_np(self._trunc).astype(bool))
        return self._host

    def __len__(self):
        return int(self._info.shape[0])

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        info, term, trunc = self._arrays()
        i = int(i)
        if i < 0:
            i += len(self)
        row, c = info[i], INFO_COLUMNS
        d = dict(offroad=float(row[c["offroad"]]), collision=float(row[c["collision"]]),
                 traffic_light_violation=float(row[c["traffic_light_violation"]]), is_success=bool(row[c["is_success"]] != 0),
                 reached_waypoint_num=int(row[c["reached_waypoint_num"]]), psi_smoothness=float(row[c["psi_smoothness"]]),
                 psi_reward=float(row[c["psi_reward"]]), dist_reward=float(row[c["dist_reward"]]),
                 speed_smoothness=float(row[c["speed_smoothness"]]), wrong_way=float(row[c["wrong_way"]]))
        if term[i] or trunc[i]:
            d["episode"] = dict(r=float(row[c["episode_return"]]), l=int(row[c["episode_length"]]), t=round(self._elapsed, 6))
            d["TimeLimit.truncated"] = bool(trunc[i] and not term[i])
            if self._tobs is not None:
                t = self._tobs[i]
                d["terminal_observation"] = t.cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
        return d


class TorchDriveVecEnv(*([_SB3VecEnv] if _SB3VecEnv is not None else [])):

    """E environments stepped in lockstep on one GPU with the SB3 ``VecEnv`` call shape
    (examples/rl_training.py:159-160: SubprocVecEnv + VecFrameStack(n_stack, channels_order="first")).

    ``reset() -> obs[E, 3*n_stack, 64, 64]``; ``step(actions[E, 2]) -> (obs, rewards[E], dones[E], infos)``;
    finished envs are re-initialised inside the step kernel (auto-reset), their frame stack restarts;
    the stack is shifted and filled by the render kernel itself (no separate roll / copy pass).
    Outputs stay on the GPU as torch tensors (``output="torch"``) or are copied to numpy (``"numpy"``).
    ``infos`` is a :class:`VecInfos`: a sequence of per-env dicts as SB3 expects (``infos[i]["terminal_observation"]``,
    ``["TimeLimit.truncated"]``, ``["episode"]`` for finished envs), materialised lazily, with the whole columns
    available as ``infos.columns``.  A subclass of SB3's ``VecEnv`` when stable-baselines3 is importable.
    ``n_stack`` defaults to ``cfg.frame_stack`` (3, as ``VecFrameStack(env, n_stack=3)`` in the reference's trainer).
    """

    def __init__(self, cfg: EnvConfig, data, num_envs: int, n_stack: Optional[int] = None, n_background: int = 0,
                 device: Optional[str] = None, output: str = "torch", env_index_offset: int = 0, seed: Optional[int] = None,
                 terminal_observation: bool = False, frame_ring: Optional[int] = None):
        self.config = cfg
        self.num_envs = int(num_envs)
        self.n_stack = int(n_stack if n_stack is not None else max(1, int(cfg.frame_stack)))
        self.output = output
        self._t_start = time.time()
        self._seed = int(seed if seed is not None else (cfg.seed if cfg.seed is not None else 0))
        self.scenario_set = data if isinstance(data, ScenarioSet) else scenario_set_from_suite(cfg, data, n_background, self._seed)
        self.engine = Engine(self.scenario_set, self.num_envs, device=device or cfg.device,
                             **engine_config(cfg, auto_reset=1, env_index_offset=int(env_index_offset)))
        self.device = self.engine.device
        self.action_space = _Box(low=np.array([-1.0, -0.3], np.float32), high=np.array([1.0, 0.3], np.float32), dtype=np.float32)
        self.observation_space = _Box(low=0, high=255, shape=(3 * self.n_stack, TDE_OBS_H, TDE_OBS_W), dtype=np.uint8)
        # The frame stack.  Default (n_stack > 1, no terminal observations): a ring of n_stack + 1 stacked observations that
        # the render kernel fills without moving a frame (tde_step_stacked_ring: the new frame goes into the observation of
        # this step and, one channel group further down each, into those of the next n_stack - 1 steps).  The tensor a step
        # returns is a slot of the ring: it stays as it is through the NEXT step (what an on-policy trainer needs, which
        # stores the observation it acted on after stepping) and is rewritten after that.  frame_ring=0 selects the single
        # tensor shifted in place by the kernel (VecFrameStack's roll-and-copy; the returned tensor is the same every step).
        ring = (self.n_stack + 1 if self.n_stack > 1 and not terminal_observation else 0) if frame_ring is None else int(frame_ring)
        if ring and (self.n_stack < 2 or terminal_observation or ring < self.n_stack):
            raise ValueError("frame_ring needs n_stack >= 2, at least n_stack slots and terminal_observation=False")
        self._ring = self.engine.new_stack_ring(self.n_stack, ring) if ring else None
        self._pos = 0
        self._stack = self._ring[0] if ring else torch.zeros((self.num_envs, 3 * self.n_stack, TDE_OBS_H, TDE_OBS_W), dtype=torch.uint8, device=self.device)
        self._actions = None
        # SB3's VecEnv keeps the last observation of a finished episode in info["terminal_observation"]; here it is
        # one [E, 3*n_stack, 64, 64] tensor whose rows are valid where `dones` is set (tde_step_terminal)
        self._terminal = torch.zeros((self.num_envs, 3 * self.n_stack, TDE_OBS_H, TDE_OBS_W), dtype=torch.uint8, device=self.device) if terminal_observation else None
        if _SB3VecEnv is not None:   # pragma: no cover
            _SB3VecEnv.__init__(self, self.num_envs, self.observation_space, self.action_space)

    def _out(self, t: torch.Tensor):
        return t.cpu().numpy() if self.output == "numpy" else t

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        if seed is not None:   # gymnasium's contract: the same seed restarts the same stream of episodes
            self._seed = int(seed)
            v = self.engine.get_env_vars()
            v[:, 5] = 0
            self.engine.set_env_vars(v)
        self._t_start = time.time()
        self.engine.reset(seed=self._seed)
        # a reset env restarts with zeros and its first frame
        if self._ring is not None:
            self._pos = 0
            self._stack = self.engine.render_stacked_ring(self._ring, self.n_stack)
        else:
            self._stack.zero_()
            self.engine.render_stacked(self._stack, self.n_stack)
        return self._out(self._stack)

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        a = torch.as_tensor(self._actions, dtype=torch.float32)
        if self._ring is not None:
            self._pos = (self._pos + 1) % self._ring.shape[0]
            self._stack, rew, term, trunc, info = self.engine.step_stacked_ring(a, self._ring, self._pos, self.n_stack)
        elif self._terminal is not None:
            _, rew, term, trunc, info = self.engine.step_terminal(a, self._stack, self._terminal, self.n_stack)
        else:
            _, rew, term, trunc, info = self.engine.step_stacked(a, self._stack, self.n_stack)
        dones = (term | trunc).bool()
        infos = VecInfos(self._out(info), self._out(term.bool()), self._out(trunc.bool()),
                         None if self._terminal is None else self._out(self._terminal), elapsed=time.time() - self._t_start)
        return self._out(self._stack), self._out(rew), self._out(dones), infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    # -- the rest of SB3's VecEnv surface (stable_baselines3.common.vec_env.base_vec_env.VecEnv): the E envs are one object
    #    here, so attribute and method calls address that object and answer once per selected env
    metadata = {"render_modes": ["rgb_array"], "render_fps": 10}
    render_mode = "rgb_array"

    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        return [int(indices)] if isinstance(indices, (int, np.integer)) else [int(i) for i in indices]

    def get_attr(self, attr_name: str, indices=None):
        """One value per selected env: row i of a per-env array / tensor attribute, the shared value otherwise."""
        v = getattr(self, attr_name)
        per_env = (torch.is_tensor(v) or isinstance(v, np.ndarray)) and v.ndim >= 1 and v.shape[0] == self.num_envs
        return [v[i] if per_env else v for i in self._indices(indices)]

    def set_attr(self, attr_name: str, value, indices=None) -> None:
        setattr(self, attr_name, value)

    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs):
        n = len(self._indices(indices))
        return [getattr(self, method_name)(*method_args, **method_kwargs)] * n

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def seed(self, seed: Optional[int] = None):
        """SB3 semantics: env i gets seed + i; here the reset RNG is keyed by (seed, global env index), so one value does it."""
        if seed is not None:
            self._seed = int(seed)
        return [self._seed + i for i in range(self.num_envs)]

    def get_images(self):
        """Last frame of every env as HWC arrays (VecEnv.get_images)."""
        frames = self._stack[:, -3:].permute(0, 2, 3, 1).cpu().numpy()
        return [f for f in frames]

    def render(self, mode: Optional[str] = None):
        """The newest birdview of env 0, H x W x 3 (rgb_array)."""
        return self._stack[0, -3:].permute(1, 2, 0).cpu().numpy()

    @property
    def unwrapped(self):
        return self

    def episode_statistics(self, reset: bool = False) -> Dict[str, float]:
        from ._capi import STAT_NAMES
        s = self.engine.episode_stats(reset=reset)
        return {n: float(s[i]) for i, n in enumerate(STAT_NAMES)}

    def close(self):
        self.engine.close()
