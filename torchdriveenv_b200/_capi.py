"""ctypes mirror of include/tde_b200.h and the loader of libtde_b200.so.

There is no CPU fallback: if the CUDA library is missing or fails to load, importing the product
path raises.  (The oracle under oracle/ is test infrastructure and is never loaded from here.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict

import numpy as np

TDE_MAX_AGENTS = 64
TDE_OBS_H = 64
TDE_OBS_W = 64
TDE_OBS_C = 3
TDE_MAX_STOPLINES = 32
TDE_INFO_STRIDE = 16
TDE_NUM_CLASSES = 11
TDE_NUM_STATS = 16

INFO_COLUMNS = {
    "offroad": 0, "collision": 1, "traffic_light_violation": 2, "is_success": 3,
    "reached_waypoint_num": 4, "psi_smoothness": 5, "psi_reward": 6, "dist_reward": 7,
    "speed_smoothness": 8, "wrong_way": 9, "episode_return": 10, "episode_length": 11,
    "scenario": 12, "did_reset": 13,
}
INFR_COLLISION, INFR_OFFROAD, INFR_TL_VIOLATION, INFR_WRONG_WAY = 0, 1, 2, 3
PH_KINEMATICS, PH_INFRACTIONS, PH_REWARD, PH_RENDER, PH_ALL = 1, 2, 4, 8, 15
STAT_NAMES = ["episodes", "return_sum", "length_sum", "offroad", "collision", "traffic_light_violation",
              "success", "reached_waypoints", "steps"]
ERROR_NAMES = {0: "TDE_OK", -1: "TDE_E_INVAL", -2: "TDE_E_CUDA", -3: "TDE_E_SHAPE", -4: "TDE_E_ARCH",
               -5: "TDE_E_STATE"}


class TdeConfig(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int32), ("max_agents", C.c_int32), ("env_index_offset", C.c_int64),
        ("max_environment_steps", C.c_int32), ("terminated_at_infraction", C.c_int32),
        ("left_handed_coordinates", C.c_int32), ("auto_reset", C.c_int32),
        ("randomize_ego_attributes", C.c_int32), ("device", C.c_int32),
        ("dt", C.c_float), ("waypoint_bonus", C.c_float), ("heading_penalty", C.c_float),
        ("distance_bonus", C.c_float), ("distance_cutoff", C.c_float), ("reach_radius", C.c_float),
        ("offroad_threshold", C.c_float), ("tl_rear_factor", C.c_float), ("fov", C.c_float),
        ("start_speed_max", C.c_float), ("start_heading_sigma", C.c_float),
        ("stage_map_tables", C.c_int32), ("host_obs_rgb", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


def default_config(**overrides) -> TdeConfig:
    """Reference defaults (EnvConfig gym_env.py:34-54; TorchDriveConfig/RendererConfig values the env
    relies on)."""
    cfg = TdeConfig()
    cfg.num_envs, cfg.max_agents, cfg.env_index_offset = 1, 1, 0
    cfg.max_environment_steps = 200
    cfg.terminated_at_infraction = 1
    cfg.left_handed_coordinates = 1
    cfg.auto_reset = 0
    cfg.randomize_ego_attributes = 0
    cfg.device = 0
    cfg.dt = 0.1
    cfg.waypoint_bonus, cfg.heading_penalty = 100.0, 25.0
    cfg.distance_bonus, cfg.distance_cutoff = 1.0, 0.5
    cfg.reach_radius = 3.0
    cfg.offroad_threshold = 0.5
    cfg.tl_rear_factor = 0.1
    cfg.fov = 35.0
    cfg.start_speed_max = 10.0
    cfg.start_heading_sigma = 0.1
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown config field {k!r}")
        setattr(cfg, k, v)
    return cfg


_I32P, _F32P, _U8P = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_uint8)


class TdeScenarioSet(C.Structure):
    _fields_ = [
        ("num_maps", C.c_int32),
        ("map_tri_offset", _I32P), ("road_tris", _F32P),
        ("map_mark_offset", _I32P), ("mark_tris", _F32P),
        ("map_stop_offset", _I32P), ("stoplines", _F32P),
        ("map_light_period", _I32P), ("map_light_offset", _I32P), ("light_states", _U8P),
        ("num_scenarios", C.c_int32),
        ("scen_map", _I32P), ("scen_wp_offset", _I32P), ("waypoints", _F32P),
        ("scen_start_heading", _F32P), ("scen_num_agents", _I32P),
        ("agent_init", _F32P), ("agent_attr", _F32P),
        ("scen_replay_T", _I32P), ("scen_replay_offset", _I32P),
        ("replay_states", _F32P), ("replay_mask", _U8P),
    ]


_PTR_TYPES = {np.dtype(np.int32): _I32P, np.dtype(np.float32): _F32P, np.dtype(np.uint8): _U8P}


def scenario_struct(packed: Dict[str, np.ndarray]):
    """Build a tde_scenario_set over the packed arrays. Returns (struct, keepalive)."""
    s = TdeScenarioSet()
    keep = []
    s.num_maps = int(packed["map_light_period"].shape[0])
    s.num_scenarios = int(packed["scen_map"].shape[0])
    for name, ctype in TdeScenarioSet._fields_:
        if name in ("num_maps", "num_scenarios"):
            continue
        arr = np.ascontiguousarray(packed[name])
        want = {_I32P: np.int32, _F32P: np.float32, _U8P: np.uint8}[ctype]
        if arr.dtype != want:
            arr = arr.astype(want)
        if arr.size == 0:  # keep a valid pointer for empty tables
            arr = np.zeros(1, want)
        keep.append(arr)
        setattr(s, name, arr.ctypes.data_as(ctype))
    return s, keep


_LIB = None
_LIB_PATH = os.environ.get("TDE_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtde_b200.so")  # the override is for A/B builds of the same CUDA library

EXPORTS = [
    "tde_version", "tde_last_error", "tde_create", "tde_destroy", "tde_default_config",
    "tde_upload_scenarios", "tde_set_env_scenario_range", "tde_set_palette", "tde_reset", "tde_step",
    "tde_step_phases", "tde_step_host", "tde_step_stacked", "tde_render_stacked", "tde_step_rollout", "tde_step_rollout_scatter", "tde_step_stacked_ring", "tde_step_terminal", "tde_kinematics", "tde_render", "tde_render_classes",
    "tde_compute_infractions",
    "tde_get_state", "tde_set_state", "tde_get_attributes", "tde_set_attributes", "tde_get_infractions",
    "tde_get_env_vars", "tde_set_env_vars", "tde_collision_boxes", "tde_offroad_boxes", "tde_clone",
    "tde_get_episode_stats", "tde_num_kernel_launches", "tde_device_sm_count", "tde_get_map_info", "tde_render_view",
]


def library_path() -> str:
    return _LIB_PATH


def load_library() -> C.CDLL:
    """Load libtde_b200.so (built in-tree by __graft_entry__.build()). Fails loudly if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). torchdriveenv_b200 has no CPU fallback.")
    lib = C.CDLL(_LIB_PATH, mode=C.RTLD_GLOBAL)
    bind_signatures(lib)
    _LIB = lib
    return lib


def bind_signatures(lib: C.CDLL) -> C.CDLL:
    """Attach the argument / result types of include/tde_b200.h to a loaded library."""
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    sig = {
        "tde_version": ([], C.c_int),
        "tde_last_error": ([vp], C.c_char_p),
        "tde_create": ([C.POINTER(TdeConfig), C.POINTER(vp)], C.c_int),
        "tde_destroy": ([vp], C.c_int),
        "tde_default_config": ([C.POINTER(TdeConfig)], C.c_int),
        "tde_upload_scenarios": ([vp, C.POINTER(TdeScenarioSet)], C.c_int),
        "tde_set_env_scenario_range": ([vp, vp, vp], C.c_int),
        "tde_set_palette": ([vp, vp], C.c_int),
        "tde_reset": ([vp, vp, u64, vp], C.c_int),
        "tde_step": ([vp, vp, vp, vp, vp, vp, vp, vp], C.c_int),
        "tde_step_phases": ([vp, i32, vp, vp, vp, vp, vp, vp, vp], C.c_int),
        "tde_step_host": ([vp, vp, vp, vp, vp, vp, vp, vp], C.c_int),
        "tde_step_stacked": ([vp, vp, vp, i32, vp, vp, vp, vp, vp], C.c_int),
        "tde_render_stacked": ([vp, vp, i32, vp], C.c_int),
        "tde_step_rollout": ([vp, vp, vp, vp, i32, vp, vp, vp, vp, vp], C.c_int),
        "tde_step_terminal": ([vp, vp, vp, i32, vp, vp, vp, vp, vp, vp], C.c_int),
        "tde_step_rollout_scatter": ([vp, vp, vp, i64, i32, i32, vp, vp, vp, vp, vp], C.c_int),
        "tde_step_stacked_ring": ([vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp], C.c_int),
        "tde_kinematics": ([vp, vp, vp], C.c_int),
        "tde_render": ([vp, vp, vp], C.c_int),
        "tde_render_classes": ([vp, vp, vp], C.c_int),
        "tde_compute_infractions": ([vp, vp], C.c_int),
        "tde_get_state": ([vp, vp, vp], C.c_int),
        "tde_set_state": ([vp, vp, vp], C.c_int),
        "tde_get_attributes": ([vp, vp, vp], C.c_int),
        "tde_set_attributes": ([vp, vp, vp], C.c_int),
        "tde_get_infractions": ([vp, vp, vp], C.c_int),
        "tde_get_env_vars": ([vp, vp, vp], C.c_int),
        "tde_set_env_vars": ([vp, vp, vp], C.c_int),
        "tde_collision_boxes": ([vp, vp, i32, i32, vp, vp], C.c_int),
        "tde_offroad_boxes": ([vp, i32, vp, vp, i32, i32, vp, vp], C.c_int),
        "tde_clone": ([vp, C.POINTER(vp), vp], C.c_int),
        "tde_get_episode_stats": ([vp, C.POINTER(C.c_double), i32, vp], C.c_int),
        "tde_num_kernel_launches": ([vp, C.POINTER(i64)], C.c_int),
        "tde_device_sm_count": ([vp, C.POINTER(i32)], C.c_int),
        "tde_get_map_info": ([vp, i32, C.POINTER(i32)], C.c_int),
        "tde_render_view": ([vp, i32, C.c_float, C.c_float, C.c_float, C.c_float, i32, i32, vp, vp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not match the header
        fn.argtypes, fn.restype = args, res
    return lib


class TdeError(RuntimeError):
    pass


def check(lib, handle, code: int, what: str):
    if code != 0:
        msg = lib.tde_last_error(handle)
        raise TdeError(f"{what}: {ERROR_NAMES.get(code, code)}: {msg.decode() if msg else ''}")
