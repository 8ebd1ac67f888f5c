"""ctypes wrapper of oracle/libtde_oracle.so — the CPU restatement of the hot path.

TEST INFRASTRUCTURE ONLY (parity unpinned, see the header of tde_oracle.c): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product.
The struct layouts are the public C ABI's (include/tde_b200.h), mirrored in
torchdriveenv_b200/_capi.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from torchdriveenv_b200._capi import (TDE_INFO_STRIDE, TDE_NUM_STATS, TDE_OBS_H, TDE_OBS_W, TdeConfig,
                                      TdeScenarioSet, scenario_struct)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtde_oracle.so")
_LIB = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tde_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        vp, f, i = C.c_void_p, C.c_float, C.c_int
        L.orc_create.argtypes, L.orc_create.restype = [C.POINTER(TdeConfig), C.POINTER(TdeScenarioSet)], vp
        L.orc_destroy.argtypes = [vp]
        L.orc_set_env_scenario_range.argtypes = [vp, vp, vp]
        L.orc_set_palette.argtypes = [vp, vp]
        L.orc_set_terminal_buffer.argtypes = [vp, vp]
        for name in ("orc_state", "orc_attr", "orc_infractions", "orc_vars", "orc_stats"):
            getattr(L, name).argtypes, getattr(L, name).restype = [vp], vp
        L.orc_reset.argtypes = [vp, vp, C.c_uint64]
        L.orc_step.argtypes = [vp] + [vp] * 6
        L.orc_step_phases.argtypes = [vp, i] + [vp] * 7
        L.orc_kinematics.argtypes = [vp, vp]
        L.orc_compute_infractions.argtypes = [vp]
        L.orc_render.argtypes = [vp, vp]
        L.orc_render_classes.argtypes = [vp, vp]
        L.orc_render_view.argtypes, L.orc_render_view.restype = [vp, i, f, f, f, f, i, i, vp], i
        L.orc_collision_boxes.argtypes = [vp, vp, i, i, vp]
        L.orc_collision_margins.argtypes = [vp, vp, i, i, vp]
        L.orc_offroad_boxes.argtypes = [vp, i, f, vp, vp, i, i, vp]
        L.orc_sincos_array.argtypes = [vp, i, vp, vp]
        L.orc_bicycle_array.argtypes = [vp, vp, vp, f, i]
        L.orc_overlap_raw.argtypes, L.orc_overlap_raw.restype = [vp, vp], i
        L.orc_point_mesh_distance.argtypes, L.orc_point_mesh_distance.restype = [vp, i, f, f], f
        L.orc_wrap_pi.argtypes, L.orc_wrap_pi.restype = [f], f
        L.orc_num_threads.restype = i
        L.orc_set_num_threads.argtypes = [i]
        L.orc_rng.argtypes, L.orc_rng.restype = [C.c_uint64] * 4, C.c_uint64
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleEnvSet:
    """E lockstep environments stepped on the CPU; mirrors the tde_* C ABI call for call."""

    def __init__(self, cfg: TdeConfig, packed: Dict[str, np.ndarray]):
        self.L = lib()
        self.cfg = cfg
        self.E, self.A = int(cfg.num_envs), int(cfg.max_agents)
        s, self._keep = scenario_struct(packed)
        self.h = self.L.orc_create(C.byref(cfg), C.byref(s))

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, fn, shape, ctype, dtype):
        ptr = C.cast(fn(self.h), C.POINTER(ctype))
        return np.ctypeslib.as_array(ptr, shape=shape).view(dtype) if False else np.ctypeslib.as_array(ptr, shape=shape)

    @property
    def state(self) -> np.ndarray:
        return self._view(self.L.orc_state, (self.E, self.A, 4), C.c_float, np.float32)

    @property
    def attr(self) -> np.ndarray:
        return self._view(self.L.orc_attr, (self.E, self.A, 4), C.c_float, np.float32)

    @property
    def infractions(self) -> np.ndarray:
        return self._view(self.L.orc_infractions, (self.E, self.A, 4), C.c_float, np.float32)

    @property
    def env_vars(self) -> np.ndarray:
        return self._view(self.L.orc_vars, (self.E, 8), C.c_int32, np.int32)

    @property
    def stats(self) -> np.ndarray:
        return self._view(self.L.orc_stats, (TDE_NUM_STATS,), C.c_double, np.float64)

    def set_env_scenario_range(self, lo, hi):
        lo = np.ascontiguousarray(lo, np.int32); hi = np.ascontiguousarray(hi, np.int32)
        self.L.orc_set_env_scenario_range(self.h, _p(lo), _p(hi))

    def set_palette(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        self.L.orc_set_palette(self.h, _p(rgb))

    def reset(self, mask: Optional[np.ndarray] = None, seed: int = 0):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self.L.orc_reset(self.h, None if m is None else _p(m), C.c_uint64(seed))

    def set_terminal_buffer(self, buf: Optional[np.ndarray]):
        """uint8[E,3,64,64] that receives the frame of an env's final state before its auto-reset (or None)."""
        self._terminal = buf
        self.L.orc_set_terminal_buffer(self.h, None if buf is None else _p(buf))

    def step(self, actions, render: bool = True, phases: int = 15):
        a = _f32(actions).reshape(self.E, 2)
        obs = np.zeros((self.E, 3, TDE_OBS_H, TDE_OBS_W), np.uint8) if render else None
        rew = np.zeros(self.E, np.float32)
        term = np.zeros(self.E, np.uint8)
        trunc = np.zeros(self.E, np.uint8)
        info = np.zeros((self.E, TDE_INFO_STRIDE), np.float32)
        self.L.orc_step_phases(self.h, int(phases), _p(a), None if obs is None else _p(obs), _p(rew),
                               _p(term), _p(trunc), _p(info), None)
        return obs, rew, term, trunc, info

    def kinematics(self, actions):
        a = _f32(actions).reshape(self.E, 2)
        self.L.orc_kinematics(self.h, _p(a))

    def compute_infractions(self) -> np.ndarray:
        self.L.orc_compute_infractions(self.h)
        return self.infractions.copy()

    def render(self) -> np.ndarray:
        obs = np.zeros((self.E, 3, TDE_OBS_H, TDE_OBS_W), np.uint8)
        self.L.orc_render(self.h, _p(obs))
        return obs

    def render_view(self, env: int, cam_x: float, cam_y: float, cam_psi: float, fov: float, width: int, height: int) -> np.ndarray:
        """uint8[3, height, width]: the recording view of one env (twin of tde_render_view)."""
        out = np.zeros((3, int(height), int(width)), np.uint8)
        rc = self.L.orc_render_view(self.h, int(env), C.c_float(cam_x), C.c_float(cam_y), C.c_float(cam_psi), C.c_float(fov),
                                    int(width), int(height), _p(out))
        if rc != 0:
            raise ValueError("orc_render_view: bad argument")
        return out

    def render_classes(self) -> np.ndarray:
        cls = np.zeros((self.E, TDE_OBS_H, TDE_OBS_W), np.uint8)
        self.L.orc_render_classes(self.h, _p(cls))
        return cls


def sincos(x):
    x = _f32(x).reshape(-1)
    s, c = np.zeros_like(x), np.zeros_like(x)
    lib().orc_sincos_array(_p(x), x.size, _p(s), _p(c))
    return s, c


def bicycle_step(state, action, lr, dt=0.1):
    st = _f32(state).reshape(-1, 4).copy()
    ac = _f32(action).reshape(-1, 2)
    l = _f32(lr).reshape(-1)
    lib().orc_bicycle_array(_p(st), _p(ac), _p(l), C.c_float(dt), st.shape[0])
    return st


def overlap(a5, b5) -> bool:
    a, b = _f32(a5), _f32(b5)
    return bool(lib().orc_overlap_raw(_p(a), _p(b)))


def collision_boxes(state, attr):
    st, at = _f32(state), _f32(attr)
    E, A = st.shape[:2]
    out = np.zeros((E, A), np.float32)
    lib().orc_collision_boxes(_p(st), _p(at), E, A, _p(out))
    return out


def collision_margins(state, attr):
    st, at = _f32(state), _f32(attr)
    E, A = st.shape[:2]
    out = np.zeros((E, A), np.float32)
    lib().orc_collision_margins(_p(st), _p(at), E, A, _p(out))
    return out


def offroad_boxes(road_tris, threshold, state, attr):
    tr, st, at = _f32(road_tris).reshape(-1, 8), _f32(state), _f32(attr)
    E, A = st.shape[:2]
    out = np.zeros((E, A), np.float32)
    lib().orc_offroad_boxes(_p(tr), tr.shape[0], C.c_float(threshold), _p(st), _p(at), E, A, _p(out))
    return out


def point_mesh_distance(road_tris, px, py) -> float:
    tr = _f32(road_tris).reshape(-1, 8)
    return float(lib().orc_point_mesh_distance(_p(tr), tr.shape[0], C.c_float(px), C.c_float(py)))


def wrap_pi(x: float) -> float:
    return float(lib().orc_wrap_pi(C.c_float(x)))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the oracle's env loops (overrides OMP_NUM_THREADS, which torchrun sets to 1)."""
    lib().orc_set_num_threads(int(n))
    return num_threads()
