"""torchdriveenv_b200 — B200-native (sm_100a) implementation of TorchDriveEnv's per-timestep
simulation hot path behind the reference's gym.Env / SimulatorInterface call surface.

Importing the package does not touch the GPU; constructing an Engine / env loads libtde_b200.so and
fails loudly when it (or a CUDA device) is missing — there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import scenarios  # noqa: F401
from ._capi import INFO_COLUMNS, default_config, library_path, load_library  # noqa: F401


def __getattr__(name):  # lazy: these need torch
    if name in ("Engine",):
        from .engine import Engine
        return Engine
    if name in ("BatchedSimulator",):
        from .simulator import BatchedSimulator
        return BatchedSimulator
    if name in ("EnvConfig", "Scenario", "WaypointSuite", "GymEnv", "WaypointSuiteEnv", "SingleAgentWrapper",
                "TorchDriveVecEnv", "build_simulator", "scenario_set_from_suite"):
        from . import gym_env
        return getattr(gym_env, name)
    raise AttributeError(name)


def register_gym():
    """gym id 'torchdriveenv-v0' taking args={'cfg','data'} (reference torchdriveenv/__init__.py:10)."""
    import gymnasium as gym
    from .gym_env import SingleAgentWrapper, WaypointSuiteEnv
    gym.register('torchdriveenv-v0', entry_point=lambda args: SingleAgentWrapper(WaypointSuiteEnv(cfg=args['cfg'], data=args['data'])))


try:  # gymnasium is optional
    register_gym()
except Exception:
    pass
