"""SimulatorInterface-level call surface over the CUDA engine.

The reference drives torchdrivesim through exactly these methods (gym_env.py): ``step`` :117,
``get_state`` :127,371,392-393, ``render_egocentric`` :123,154, ``compute_offroad`` :142,415,427,
``compute_collision`` :143,415,428, ``compute_traffic_lights_violations`` :144,415,429, ``to`` :298,
``copy`` :110 (plus ``compute_wrong_way`` of the same interface, never called by the env).
Tensors are ``B x A x ...`` with B = lockstep envs and, because NPCs are hidden from the interface
exactly as ``IAIWrapper`` hides them (npc_mask :269-271), A = 1 (the ego) unless ``expose_npcs``.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._capi import INFR_COLLISION, INFR_OFFROAD, INFR_TL_VIOLATION, INFR_WRONG_WAY, PH_INFRACTIONS, PH_KINEMATICS
from .engine import Engine
from .scenarios import ScenarioSet


class BatchedSimulator:
    def __init__(self, scenarios: ScenarioSet, num_envs: int = 1, max_agents: Optional[int] = None,
                 device: Optional[str] = None, expose_npcs: bool = False, seed: int = 0, **config):
        self._args = dict(scenarios=scenarios, num_envs=num_envs, max_agents=max_agents, device=device,
                          expose_npcs=expose_npcs, seed=seed, config=dict(config))
        self.engine = Engine(scenarios, num_envs, max_agents, device=device, **config)
        self.expose_npcs = expose_npcs
        self.seed = seed
        self.engine.reset(seed=seed)
        self._infractions_valid = False

    # -- SimulatorInterface
    @property
    def device(self):
        return self.engine.device

    @property
    def batch_size(self) -> int:
        return self.engine.E

    def _agents(self, t: torch.Tensor) -> torch.Tensor:
        return t if self.expose_npcs else t[:, :1]

    def step(self, action: torch.Tensor) -> None:
        """action: B x A x 2 (acceleration, steering) in physical units; only the ego's row is used,
        NPCs follow their log-replay / constant-velocity drive."""
        a = torch.as_tensor(action, dtype=torch.float32)
        a = a.reshape(self.engine.E, -1, 2)[:, 0]
        # kinematics + infractions in one launch; reward/termination belong to the env layer
        self.engine.step(a, render=False, phases=PH_KINEMATICS | PH_INFRACTIONS)
        self._infractions_valid = True

    def get_state(self) -> torch.Tensor:
        return self._agents(self.engine.get_state())

    def set_state(self, state: torch.Tensor) -> None:
        full = self.engine.get_state()
        state = torch.as_tensor(state, dtype=torch.float32, device=self.device)
        if state.shape[1] == full.shape[1]:
            full = state
        else:
            full[:, : state.shape[1]] = state
        self.engine.set_state(full)
        self._infractions_valid = False

    def get_agent_size(self) -> torch.Tensor:
        return self._agents(self.engine.get_attributes()[..., :2])

    def get_present_mask(self) -> torch.Tensor:
        return self._agents(self.engine.get_attributes()[..., 3] != 0)

    def render_egocentric(self) -> torch.Tensor:
        """B x 1 x 3 x 64 x 64 uint8 birdview centred on each env's ego."""
        return self.engine.render().unsqueeze(1)

    def _infractions(self) -> torch.Tensor:
        if not self._infractions_valid:
            self.engine.compute_infractions()
            self._infractions_valid = True
        return self.engine.get_infractions()

    def compute_offroad(self) -> torch.Tensor:
        return self._agents(self._infractions()[..., INFR_OFFROAD])

    def compute_collision(self) -> torch.Tensor:
        return self._agents(self._infractions()[..., INFR_COLLISION])

    def compute_traffic_lights_violations(self) -> torch.Tensor:
        return self._agents(self._infractions()[..., INFR_TL_VIOLATION])

    def compute_wrong_way(self) -> torch.Tensor:
        return self._agents(self._infractions()[..., INFR_WRONG_WAY])

    def to(self, device) -> "BatchedSimulator":
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("BatchedSimulator.to(): only CUDA devices are supported, there is no CPU path")
        if device.index is not None and device.index != self.device.index:
            a = dict(self._args); a["device"] = str(device)
            other = BatchedSimulator(a["scenarios"], a["num_envs"], a["max_agents"], a["device"], a["expose_npcs"], a["seed"], **a["config"])
            other._copy_from(self)
            return other
        return self

    def copy(self) -> "BatchedSimulator":
        """simulator.copy() :110 -> tde_clone: device-to-device copy of every env, scenario tables shared."""
        other = object.__new__(BatchedSimulator)
        other._args = self._args
        other.engine = self.engine.clone()
        other.expose_npcs, other.seed = self.expose_npcs, self.seed
        other._infractions_valid = self._infractions_valid
        return other

    def _copy_from(self, src: "BatchedSimulator") -> None:
        self.engine.set_state(src.engine.get_state().to(self.device))
        self.engine.set_attributes(src.engine.get_attributes().to(self.device))
        self.engine.set_env_vars(src.engine.get_env_vars().to(self.device))
        self._infractions_valid = False


class BirdviewRecordingWrapper:
    """Records a large top-down frame of env 0 after construction and after every step, as the reference's
    ``BirdviewRecordingWrapper(simulator, res=Resolution(video_res, video_res), fov=video_fov, to_cpu=True)``
    does (gym_env.py:52-53, :295-297); ``get_birdviews()`` (:174) hands the frames to ``save_video``.
    Everything else is forwarded to the wrapped simulator.  The frames come from ``tde_render_view``
    (camera fixed on the centre of the map's road mesh unless ``camera_xy`` is given)."""

    def __init__(self, simulator: BatchedSimulator, res=(1024, 1024), fov: float = 500.0, camera_xy=None,
                 camera_psi: float = 0.0, to_cpu: bool = True, env: int = 0):
        self.simulator = simulator
        self.res = (int(res[0]), int(res[1])) if not isinstance(res, int) else (int(res), int(res))
        self.fov, self.camera_xy, self.camera_psi, self.to_cpu, self.env = float(fov), camera_xy, float(camera_psi), bool(to_cpu), int(env)
        self.birdviews = []
        self._record()

    def _record(self):
        frame = self.simulator.engine.render_view(self.env, self.camera_xy, self.camera_psi, self.fov, self.res).unsqueeze(0)
        self.birdviews.append(frame.cpu() if self.to_cpu else frame.clone())

    def step(self, action):
        self.simulator.step(action)
        self._record()

    def get_birdviews(self):
        return self.birdviews

    def copy(self):
        other = BirdviewRecordingWrapper.__new__(BirdviewRecordingWrapper)
        other.__dict__.update(self.__dict__)
        other.simulator = self.simulator.copy()
        other.birdviews = list(self.birdviews)
        return other

    def to(self, device):
        self.simulator = self.simulator.to(device)
        return self

    def __getattr__(self, name):
        return getattr(self.simulator, name)
