"""EmuEngine — TEST INFRASTRUCTURE: drives tests/emu/libtde_emu.so, i.e. the product's CUDA sources compiled
for the host-side lockstep emulator (tests/emu/cuda_emu.h), through the same C ABI with numpy buffers.

It lets the kernels' logic be checked against the oracle in a container without a GPU.  The product never
loads this library (torchdriveenv_b200 has no CPU path); the `-m gpu` tests remain the parity tests proper."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from torchdriveenv_b200 import _capi
from torchdriveenv_b200._capi import (PH_ALL, PH_RENDER, TDE_INFO_STRIDE, TDE_NUM_STATS, TDE_OBS_C, TDE_OBS_H, TDE_OBS_W,
                                      check, default_config, scenario_struct)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "emu", "libtde_emu.so")
_LIB = None


def build(force: bool = False) -> str:
    csrc = os.path.join(os.path.dirname(_HERE), "torchdriveenv_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(_HERE, "emu", "cuda_emu.h"),
                                                                os.path.join(os.path.dirname(_HERE), "include", "tde_b200.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call([os.path.join(_HERE, "emu", "build.sh")], stdout=subprocess.DEVNULL)
    return _SO


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        build()
        _LIB = _capi.bind_signatures(C.CDLL(_SO))
    return _LIB


def _p(a: Optional[np.ndarray]):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _aligned(shape, dtype) -> np.ndarray:
    """zero-initialised array whose data pointer is 16-byte aligned (the kernels move rows with 128-bit accesses)"""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + 16, np.uint8)
    off = (-raw.ctypes.data) % 16
    return raw[off:off + n].view(dtype).reshape(shape)


class EmuEngine:
    """Mirror of torchdriveenv_b200.engine.Engine over the emulated library; tensors are numpy arrays."""

    def __init__(self, scenarios, num_envs: int, max_agents: Optional[int] = None, **config):
        self.lib = lib()
        A = int(max_agents if max_agents is not None else scenarios.max_agents())
        self.E, self.A = int(num_envs), A
        self.cfg = default_config(num_envs=self.E, max_agents=A, **config)
        self.packed: Dict[str, np.ndarray] = scenarios.pack(A)
        self.scenarios = scenarios
        self.h = C.c_void_p()
        check(self.lib, None, self.lib.tde_create(C.byref(self.cfg), C.byref(self.h)), "tde_create")
        s, keep = scenario_struct(self.packed)
        check(self.lib, self.h, self.lib.tde_upload_scenarios(self.h, C.byref(s)), "tde_upload_scenarios")
        E = self.E
        self.obs = _aligned((E, TDE_OBS_C, TDE_OBS_H, TDE_OBS_W), np.uint8)
        self.reward = _aligned((E,), np.float32)
        self.terminated = _aligned((E,), np.uint8)
        self.truncated = _aligned((E,), np.uint8)
        self.info = _aligned((E, TDE_INFO_STRIDE), np.float32)

    def close(self):
        if getattr(self, "h", None):
            self.lib.tde_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, code, what):
        check(self.lib, self.h, code, what)

    def _act(self, actions):
        a = _aligned((self.E, 2), np.float32)
        a[...] = np.asarray(actions, np.float32).reshape(self.E, 2)
        return a

    def set_env_scenario_range(self, lo, hi):
        lo = np.ascontiguousarray(lo, np.int32); hi = np.ascontiguousarray(hi, np.int32)
        self._check(self.lib.tde_set_env_scenario_range(self.h, _p(lo), _p(hi)), "tde_set_env_scenario_range")

    def set_palette(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        self._check(self.lib.tde_set_palette(self.h, _p(rgb)), "tde_set_palette")

    def reset(self, mask=None, seed: int = 0):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._check(self.lib.tde_reset(self.h, _p(m), C.c_uint64(seed), None), "tde_reset")

    def step(self, actions, render: bool = True, phases: int = PH_ALL):
        a = self._act(actions)
        if not render:
            phases &= ~PH_RENDER
        self._check(self.lib.tde_step_phases(self.h, int(phases), _p(a), _p(self.obs) if render else None, _p(self.reward),
                                             _p(self.terminated), _p(self.truncated), _p(self.info), None), "tde_step")
        return (self.obs if render else None), self.reward, self.terminated, self.truncated, self.info

    def new_stack(self, n_stack: int) -> np.ndarray:
        return _aligned((self.E, 3 * n_stack, TDE_OBS_H, TDE_OBS_W), np.uint8)

    def step_stacked(self, actions, stack, n_stack: int):
        a = self._act(actions)
        self._check(self.lib.tde_step_stacked(self.h, _p(a), _p(stack), int(n_stack), _p(self.reward), _p(self.terminated),
                                              _p(self.truncated), _p(self.info), None), "tde_step_stacked")
        return stack, self.reward, self.terminated, self.truncated, self.info

    def step_terminal(self, actions, obs, terminal_obs, n_stack: int = 1):
        a = self._act(actions)
        self._check(self.lib.tde_step_terminal(self.h, _p(a), _p(obs), int(n_stack), _p(terminal_obs), _p(self.reward),
                                               _p(self.terminated), _p(self.truncated), _p(self.info), None), "tde_step_terminal")
        return obs, self.reward, self.terminated, self.truncated, self.info

    def step_rollout(self, actions, stack_prev, stack_next, n_stack: int):
        a = self._act(actions)
        self._check(self.lib.tde_step_rollout(self.h, _p(a), _p(stack_prev), _p(stack_next), int(n_stack), _p(self.reward),
                                              _p(self.terminated), _p(self.truncated), _p(self.info), None), "tde_step_rollout")
        return stack_next, self.reward, self.terminated, self.truncated, self.info

    def step_rollout_scatter(self, actions, buffer_obs, t: int, n_stack: int):
        a = self._act(actions)
        stride = buffer_obs.strides[0]
        ahead = min(int(n_stack), buffer_obs.shape[0] - 1 - t)
        self._check(self.lib.tde_step_rollout_scatter(self.h, _p(a), _p(buffer_obs[t + 1]), stride, ahead, int(n_stack), _p(self.reward),
                                                      _p(self.terminated), _p(self.truncated), _p(self.info), None), "tde_step_rollout_scatter")
        return buffer_obs[t + 1], self.reward, self.terminated, self.truncated, self.info

    def step_stacked_ring(self, actions, ring, pos: int, n_stack: int):
        a = self._act(actions)
        R = ring.shape[0]
        self._check(self.lib.tde_step_stacked_ring(self.h, _p(a), _p(ring), int(R), int(pos) % R, int(n_stack), _p(self.reward),
                                                   _p(self.terminated), _p(self.truncated), _p(self.info), None), "tde_step_stacked_ring")
        return ring[int(pos) % R], self.reward, self.terminated, self.truncated, self.info

    def render_stacked(self, stack, n_stack: int):
        self._check(self.lib.tde_render_stacked(self.h, _p(stack), int(n_stack), None), "tde_render_stacked")
        return stack

    def step_host(self, actions, render: bool = True):
        a = self._act(actions)
        self._check(self.lib.tde_step_host(self.h, _p(a), _p(self.obs) if render else None, _p(self.reward), _p(self.terminated),
                                           _p(self.truncated), _p(self.info), None), "tde_step_host")
        return (self.obs if render else None), self.reward, self.terminated, self.truncated, self.info

    def kinematics(self, actions):
        self._check(self.lib.tde_kinematics(self.h, _p(self._act(actions)), None), "tde_kinematics")

    def render(self, out=None):
        out = self.obs if out is None else out
        self._check(self.lib.tde_render(self.h, _p(out), None), "tde_render")
        return out

    def render_classes(self):
        nib = _aligned((self.E, TDE_OBS_H, TDE_OBS_W // 2), np.uint8)
        self._check(self.lib.tde_render_classes(self.h, _p(nib), None), "tde_render_classes")
        return np.stack((nib & 15, nib >> 4), axis=-1).reshape(self.E, TDE_OBS_H, TDE_OBS_W)

    def render_view(self, env, cam_x, cam_y, cam_psi, fov, width, height):
        out = _aligned((3, int(height), int(width)), np.uint8)
        self._check(self.lib.tde_render_view(self.h, int(env), float(cam_x), float(cam_y), float(cam_psi), float(fov), int(width),
                                             int(height), _p(out), None), "tde_render_view")
        return out

    def compute_infractions(self):
        self._check(self.lib.tde_compute_infractions(self.h, None), "tde_compute_infractions")
        return self.get_infractions()

    def _get(self, fn, shape, dtype, what):
        out = _aligned(shape, dtype)
        self._check(fn(self.h, _p(out), None), what)
        return out

    def _set(self, fn, t, shape, dtype, what):
        a = _aligned(shape, dtype)
        a[...] = np.asarray(t, dtype).reshape(shape)
        self._check(fn(self.h, _p(a), None), what)

    def get_state(self): return self._get(self.lib.tde_get_state, (self.E, self.A, 4), np.float32, "tde_get_state")
    def set_state(self, t): self._set(self.lib.tde_set_state, t, (self.E, self.A, 4), np.float32, "tde_set_state")
    def get_attributes(self): return self._get(self.lib.tde_get_attributes, (self.E, self.A, 4), np.float32, "tde_get_attributes")
    def set_attributes(self, t): self._set(self.lib.tde_set_attributes, t, (self.E, self.A, 4), np.float32, "tde_set_attributes")
    def get_infractions(self): return self._get(self.lib.tde_get_infractions, (self.E, self.A, 4), np.float32, "tde_get_infractions")
    def get_env_vars(self): return self._get(self.lib.tde_get_env_vars, (self.E, 8), np.int32, "tde_get_env_vars")
    def set_env_vars(self, t): self._set(self.lib.tde_set_env_vars, t, (self.E, 8), np.int32, "tde_set_env_vars")

    def episode_stats(self, reset: bool = False) -> np.ndarray:
        out = (C.c_double * TDE_NUM_STATS)()
        self._check(self.lib.tde_get_episode_stats(self.h, out, int(reset), None), "tde_get_episode_stats")
        return np.asarray(list(out), np.float64)

    def collision_boxes(self, state, attr):
        E, A = state.shape[:2]
        st, at = _aligned((E, A, 4), np.float32), _aligned((E, A, 4), np.float32)
        st[...] = state; at[...] = attr
        out = _aligned((E, A), np.float32)
        check(self.lib, None, self.lib.tde_collision_boxes(_p(st), _p(at), E, A, _p(out), None), "tde_collision_boxes")
        return out

    def offroad_boxes(self, map_id, state, attr):
        E, A = state.shape[:2]
        st, at = _aligned((E, A, 4), np.float32), _aligned((E, A, 4), np.float32)
        st[...] = state; at[...] = attr
        out = _aligned((E, A), np.float32)
        self._check(self.lib.tde_offroad_boxes(self.h, int(map_id), _p(st), _p(at), E, A, _p(out), None), "tde_offroad_boxes")
        return out
