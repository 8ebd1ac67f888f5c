"""Thin Python host over the C ABI (include/tde_b200.h): owns the handle, allocates I/O tensors with
PyTorch (device memory + streams only) and passes raw pointers to libtde_b200.so through ctypes."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _capi
from ._capi import (INFO_COLUMNS, PH_ALL, PH_RENDER, TDE_INFO_STRIDE, TDE_NUM_STATS, TDE_OBS_C, TDE_OBS_H,
                    TDE_OBS_W, check, default_config, load_library, scenario_struct)
from .scenarios import ScenarioSet


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    """E lockstep environments on one GPU.  All tensors returned live on that GPU."""

    def __init__(self, scenarios: ScenarioSet, num_envs: int, max_agents: Optional[int] = None,
                 device: Optional[str] = None, **config):
        if not torch.cuda.is_available():
            raise RuntimeError("torchdriveenv_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")
        self.lib = load_library()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError(f"device {self.device} is not a CUDA device: there is no CPU fallback")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self.scenarios = scenarios
        A = int(max_agents if max_agents is not None else scenarios.max_agents())
        self.E, self.A = int(num_envs), A
        self.cfg = default_config(num_envs=self.E, max_agents=A, device=dev_index, **config)
        self.packed: Dict[str, np.ndarray] = scenarios.pack(A)
        self.h = C.c_void_p()
        check(self.lib, None, self.lib.tde_create(C.byref(self.cfg), C.byref(self.h)), "tde_create")
        s, keep = scenario_struct(self.packed)
        check(self.lib, self.h, self.lib.tde_upload_scenarios(self.h, C.byref(s)), "tde_upload_scenarios")
        del keep
        E = self.E
        with torch.cuda.device(self.device):
            self.obs = torch.zeros((E, TDE_OBS_C, TDE_OBS_H, TDE_OBS_W), dtype=torch.uint8, device=self.device)
            self.reward = torch.zeros(E, dtype=torch.float32, device=self.device)
            self.terminated = torch.zeros(E, dtype=torch.uint8, device=self.device)
            self.truncated = torch.zeros(E, dtype=torch.uint8, device=self.device)
            self.info = torch.zeros((E, TDE_INFO_STRIDE), dtype=torch.float32, device=self.device)
        self._host = None

    # -- lifetime
    def clone(self) -> "Engine":
        """simulator.copy() (gym_env.py:110) at the ABI: tde_clone - an independent engine on the same GPU with a copy of
        every env (device-to-device), sharing the read-only scenario tables with this one."""
        other = object.__new__(Engine)
        other.lib, other.device, other.scenarios = self.lib, self.device, self.scenarios
        other.E, other.A, other.cfg, other.packed = self.E, self.A, self.cfg, self.packed
        other.h = C.c_void_p()
        with torch.cuda.device(self.device):
            self._check(self.lib.tde_clone(self.h, C.byref(other.h), self._stream()), "tde_clone")
            other.obs = self.obs.clone()
            other.reward, other.terminated, other.truncated = self.reward.clone(), self.terminated.clone(), self.truncated.clone()
            other.info = self.info.clone()
        other._host = None
        return other

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.tde_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, code, what):
        check(self.lib, self.h, code, what)

    # -- configuration
    def set_env_scenario_range(self, lo, hi):
        lo = np.ascontiguousarray(lo, np.int32); hi = np.ascontiguousarray(hi, np.int32)
        self._check(self.lib.tde_set_env_scenario_range(self.h, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p)),
                    "tde_set_env_scenario_range")

    def set_palette(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        self._check(self.lib.tde_set_palette(self.h, rgb.ctypes.data_as(C.c_void_p)), "tde_set_palette")

    # -- the hot path
    def reset(self, mask: Optional[torch.Tensor] = None, seed: int = 0):
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        self._check(self.lib.tde_reset(self.h, _ptr(mask), C.c_uint64(seed), self._stream()), "tde_reset")

    def step(self, actions: torch.Tensor, render: bool = True, phases: int = PH_ALL):
        a = actions.to(device=self.device, dtype=torch.float32).contiguous().view(self.E, 2)
        if not render:
            phases &= ~PH_RENDER
        self._check(self.lib.tde_step_phases(self.h, int(phases), _ptr(a), _ptr(self.obs) if render else None,
                                             _ptr(self.reward), _ptr(self.terminated), _ptr(self.truncated),
                                             _ptr(self.info), self._stream()), "tde_step")
        return (self.obs if render else None), self.reward, self.terminated, self.truncated, self.info

    def capture_steps(self, actions_seq: torch.Tensor, render: bool = True) -> "torch.cuda.CUDAGraph":
        """One CUDA graph of ``len(actions_seq)`` consecutive steps (actions_seq: float32[G, E, 2] on this device, read at
        every replay).  A small batch steps in about 10 us, less than it takes the host to issue a launch: replaying a graph
        keeps the GPU fed (tde_step only enqueues on the caller's stream, so it captures).  After ``graph.replay()`` the
        engine's output tensors hold the results of the last step of the block."""
        seq = actions_seq.to(device=self.device, dtype=torch.float32).contiguous()
        if seq.dim() != 3 or tuple(seq.shape[1:]) != (self.E, 2):
            raise ValueError(f"actions_seq must have shape [G, {self.E}, 2]")
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        # thread_local: CUDA calls of other threads (a NCCL watchdog, a clock sampler) must not abort the capture
        with torch.cuda.device(self.device), torch.cuda.graph(graph, capture_error_mode="thread_local"):
            for j in range(seq.shape[0]):
                self.step(seq[j], render=render)
        graph._tde_actions = seq      # the graph reads this tensor at every replay: keep it alive with the graph
        return graph

    def step_into(self, actions: torch.Tensor, obs: torch.Tensor, reward: Optional[torch.Tensor] = None,
                  terminated: Optional[torch.Tensor] = None, truncated: Optional[torch.Tensor] = None,
                  info: Optional[torch.Tensor] = None):
        """tde_step with the caller's rows: the new frame goes to `obs` uint8[E, 3, 64, 64] (e.g. one slot of a ring of
        frames), the per-env results to the given rows (contiguous, on this device) or to the engine's own tensors."""
        a = actions.to(device=self.device, dtype=torch.float32).contiguous().view(self.E, 2)
        self._check_stack(obs, 1)
        outs = []
        for x, own, shape, dt in ((reward, self.reward, (self.E,), torch.float32), (terminated, self.terminated, (self.E,), torch.uint8),
                                  (truncated, self.truncated, (self.E,), torch.uint8), (info, self.info, (self.E, TDE_INFO_STRIDE), torch.float32)):
            if x is None:
                x = own
            elif tuple(x.shape) != shape or x.dtype != dt or not x.is_contiguous() or x.device != self.obs.device:
                raise ValueError(f"step_into: output rows must be contiguous {dt} tensors of shape {shape} on {self.obs.device}")
            outs.append(x)
        self._check(self.lib.tde_step(self.h, _ptr(a), _ptr(obs), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(outs[3]),
                                      self._stream()), "tde_step")
        return obs, outs[0], outs[1], outs[2], outs[3]

    def _check_stack(self, stack: torch.Tensor, n_stack: int):
        want = (self.E, 3 * n_stack, TDE_OBS_H, TDE_OBS_W)
        if tuple(stack.shape) != want or stack.dtype != torch.uint8 or not stack.is_contiguous() or stack.device != self.obs.device:
            raise ValueError(f"stack must be a contiguous uint8 tensor of shape {want} on {self.obs.device}")

    def step_stacked(self, actions: torch.Tensor, stack: torch.Tensor, n_stack: int):
        """tde_step with VecFrameStack fused into the observation store: `stack` [E, 3*n_stack, 64, 64]
        is shifted by one frame and receives the new frame in its last three channels, in place."""
        a = actions.to(device=self.device, dtype=torch.float32).contiguous().view(self.E, 2)
        self._check_stack(stack, n_stack)
        self._check(self.lib.tde_step_stacked(self.h, _ptr(a), _ptr(stack), int(n_stack), _ptr(self.reward), _ptr(self.terminated),
                                              _ptr(self.truncated), _ptr(self.info), self._stream()), "tde_step_stacked")
        return stack, self.reward, self.terminated, self.truncated, self.info

    def step_terminal(self, actions: torch.Tensor, obs: torch.Tensor, terminal_obs: torch.Tensor, n_stack: int = 1):
        """tde_step_terminal: the (stacked) step that keeps the terminal observation of finished envs in
        `terminal_obs` (same shape as `obs`; only the rows of envs that finished in this step are written)."""
        a = actions.to(device=self.device, dtype=torch.float32).contiguous().view(self.E, 2)
        self._check_stack(obs, n_stack)
        self._check_stack(terminal_obs, n_stack)
        self._check(self.lib.tde_step_terminal(self.h, _ptr(a), _ptr(obs), int(n_stack), _ptr(terminal_obs), _ptr(self.reward),
                                               _ptr(self.terminated), _ptr(self.truncated), _ptr(self.info), self._stream()),
                    "tde_step_terminal")
        return obs, self.reward, self.terminated, self.truncated, self.info

    def step_rollout(self, actions: torch.Tensor, stack_prev: torch.Tensor, stack_next: torch.Tensor, n_stack: int,
                     reward: Optional[torch.Tensor] = None, terminated: Optional[torch.Tensor] = None,
                     truncated: Optional[torch.Tensor] = None, info: Optional[torch.Tensor] = None):
        """tde_step_rollout: the stacked step out of place - older frames read from `stack_prev` (slot t of a
        rollout buffer), shifted stack + new frame written to `stack_next` (slot t + 1).  The per-env results
        go to the given rows (contiguous, on this device) or to the engine's own output tensors."""
        a = actions.to(device=self.device, dtype=torch.float32).contiguous().view(self.E, 2)
        self._check_stack(stack_prev, n_stack)
        self._check_stack(stack_next, n_stack)
        outs = []
        for t, own, shape, dt in ((reward, self.reward, (self.E,), torch.float32), (terminated, self.terminated, (self.E,), torch.uint8),
                                  (truncated, self.truncated, (self.E,), torch.uint8), (info, self.info, (self.E, TDE_INFO_STRIDE), torch.float32)):
            if t is None:
                t = own
            elif tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or t.device != self.obs.device:
                raise ValueError(f"step_rollout: output rows must be contiguous {dt} tensors of shape {shape} on {self.obs.device}")
            outs.append(t)
        self._check(self.lib.tde_step_rollout(self.h, _ptr(a), _ptr(stack_prev), _ptr(stack_next), int(n_stack), _ptr(outs[0]),
                                              _ptr(outs[1]), _ptr(outs[2]), _ptr(outs[3]), self._stream()), "tde_step_rollout")
        return stack_next, outs[0], outs[1], outs[2], outs[3]

    def step_rollout_scatter(self, actions: torch.Tensor, buffer_obs: torch.Tensor, t: int, n_stack: int,
                             reward: Optional[torch.Tensor] = None, terminated: Optional[torch.Tensor] = None,
                             truncated: Optional[torch.Tensor] = None, info: Optional[torch.Tensor] = None):
        """tde_step_rollout_scatter: step t of a rollout whose stacked observations live in ``buffer_obs``
        [T + 1, E, 3*n_stack, 64, 64] (contiguous).  The new frame is stored into slot t + 1 and, one channel group further
        down each, into the following n_stack - 1 slots (as far as the buffer goes); nothing is read or moved."""
        a = actions.to(device=self.device, dtype=torch.float32).contiguous().view(self.E, 2)
        T1 = int(buffer_obs.shape[0])
        if buffer_obs.dim() != 5 or not buffer_obs.is_contiguous() or buffer_obs.dtype != torch.uint8 or buffer_obs.device != self.obs.device:
            raise ValueError("step_rollout_scatter: buffer_obs must be a contiguous uint8 [T + 1, E, 3*n_stack, 64, 64] tensor on the engine's device")
        self._check_stack(buffer_obs[0], n_stack)
        if not 0 <= t < T1 - 1:
            raise ValueError("step_rollout_scatter: t out of range")
        outs = []
        for x, own, shape, dt in ((reward, self.reward, (self.E,), torch.float32), (terminated, self.terminated, (self.E,), torch.uint8),
                                  (truncated, self.truncated, (self.E,), torch.uint8), (info, self.info, (self.E, TDE_INFO_STRIDE), torch.float32)):
            if x is None:
                x = own
            elif tuple(x.shape) != shape or x.dtype != dt or not x.is_contiguous() or x.device != self.obs.device:
                raise ValueError(f"step_rollout_scatter: output rows must be contiguous {dt} tensors of shape {shape} on {self.obs.device}")
            outs.append(x)
        stride = int(buffer_obs.stride(0)) * buffer_obs.element_size()
        ahead = min(int(n_stack), T1 - 1 - t)
        self._check(self.lib.tde_step_rollout_scatter(self.h, _ptr(a), _ptr(buffer_obs[t + 1]), stride, ahead, int(n_stack), _ptr(outs[0]),
                                                      _ptr(outs[1]), _ptr(outs[2]), _ptr(outs[3]), self._stream()), "tde_step_rollout_scatter")
        return buffer_obs[t + 1], outs[0], outs[1], outs[2], outs[3]

    def new_stack_ring(self, n_stack: int, ring_slots: Optional[int] = None) -> torch.Tensor:
        """uint8[ring_slots, E, 3*n_stack, 64, 64] for step_stacked_ring (default: n_stack + 1 slots)."""
        R = int(ring_slots if ring_slots is not None else n_stack + 1)
        return torch.zeros((R, self.E, 3 * int(n_stack), TDE_OBS_H, TDE_OBS_W), dtype=torch.uint8, device=self.device)

    def render_stacked_ring(self, ring: torch.Tensor, n_stack: int) -> torch.Tensor:
        """The reset observation (zeros + the first frame) into slot 0 of the ring, and its frames one channel group further
        down into the following n_stack - 1 slots, as step_stacked_ring keeps them from then on."""
        self._check_stack(ring[0], n_stack)
        ring[0].zero_()
        self.render_stacked(ring[0], n_stack)
        for j in range(1, min(int(n_stack), ring.shape[0])):
            ring[j][:, : 3 * (n_stack - j)].copy_(ring[0][:, 3 * j:])
        return ring[0]

    def step_stacked_ring(self, actions: torch.Tensor, ring: torch.Tensor, pos: int, n_stack: int):
        """tde_step_stacked_ring: VecFrameStack without moving a frame.  The step's stacked observation is ``ring[pos]``; the
        caller advances pos by one (modulo the ring size) per step.  ``ring[pos]`` stays intact for
        ``ring.shape[0] - n_stack`` further steps."""
        a = actions.to(device=self.device, dtype=torch.float32).contiguous().view(self.E, 2)
        if ring.dim() != 5 or not ring.is_contiguous() or ring.dtype != torch.uint8 or ring.device != self.obs.device:
            raise ValueError("step_stacked_ring: ring must be a contiguous uint8 [R, E, 3*n_stack, 64, 64] tensor on the engine's device")
        self._check_stack(ring[0], n_stack)
        R = int(ring.shape[0])
        self._check(self.lib.tde_step_stacked_ring(self.h, _ptr(a), _ptr(ring), R, int(pos) % R, int(n_stack), _ptr(self.reward),
                                                   _ptr(self.terminated), _ptr(self.truncated), _ptr(self.info), self._stream()),
                    "tde_step_stacked_ring")
        return ring[int(pos) % R], self.reward, self.terminated, self.truncated, self.info

    def render_stacked(self, stack: torch.Tensor, n_stack: int) -> torch.Tensor:
        self._check_stack(stack, n_stack)
        self._check(self.lib.tde_render_stacked(self.h, _ptr(stack), int(n_stack), self._stream()), "tde_render_stacked")
        return stack

    def step_host(self, actions: np.ndarray, render: bool = True):
        """Reference-facing call with HOST buffers (tde_step_host): H2D, step, D2H, synchronise."""
        E = self.E
        if self._host is None:
            pin = dict(pin_memory=True)
            self._host = dict(
                act=torch.zeros((E, 2), dtype=torch.float32, **pin),
                obs=torch.zeros((E, TDE_OBS_C, TDE_OBS_H, TDE_OBS_W), dtype=torch.uint8, **pin),
                rew=torch.zeros(E, dtype=torch.float32, **pin), term=torch.zeros(E, dtype=torch.uint8, **pin),
                trunc=torch.zeros(E, dtype=torch.uint8, **pin),
                info=torch.zeros((E, TDE_INFO_STRIDE), dtype=torch.float32, **pin))
        hb = self._host
        hb["act"].numpy()[...] = np.asarray(actions, np.float32).reshape(E, 2)
        self._check(self.lib.tde_step_host(self.h, _ptr(hb["act"]), _ptr(hb["obs"]) if render else None, _ptr(hb["rew"]),
                                           _ptr(hb["term"]), _ptr(hb["trunc"]), _ptr(hb["info"]), self._stream()),
                    "tde_step_host")
        return (hb["obs"].numpy() if render else None), hb["rew"].numpy(), hb["term"].numpy(), hb["trunc"].numpy(), hb["info"].numpy()

    def kinematics(self, actions: torch.Tensor):
        a = actions.to(device=self.device, dtype=torch.float32).contiguous().view(self.E, 2)
        self._check(self.lib.tde_kinematics(self.h, _ptr(a), self._stream()), "tde_kinematics")

    def render(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = self.obs if out is None else out
        self._check(self.lib.tde_render(self.h, _ptr(out), self._stream()), "tde_render")
        return out

    def render_classes(self, out: Optional[torch.Tensor] = None, unpack: bool = True) -> torch.Tensor:
        """tde_render_classes: the class-index birdview.  The kernel writes uint8[E, 64, 32] (4 bits per pixel, even
        pixel in the low nibble); unpack=True returns uint8[E, 64, 64] class indices (two torch ops on that result)."""
        if out is None:
            out = torch.empty((self.E, TDE_OBS_H, TDE_OBS_W // 2), dtype=torch.uint8, device=self.device)
        self._check(self.lib.tde_render_classes(self.h, _ptr(out), self._stream()), "tde_render_classes")
        if not unpack:
            return out
        return torch.stack((out & 15, out >> 4), dim=-1).reshape(self.E, TDE_OBS_H, TDE_OBS_W)

    def render_view(self, env: int = 0, camera_xy=None, camera_psi: float = 0.0, fov: float = 500.0, res=(1024, 1024),
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Recording view (tde_render_view): env `env` as a uint8[3, H, W] frame from a free camera.  camera_xy = None
        centres the camera on the road mesh of the env's map (what a fixed whole-map recording camera shows)."""
        W, H = int(res[0]), int(res[1])
        if camera_xy is None:
            m = int(self.get_env_vars()[int(env), 6])
            tris = np.asarray(self.scenarios.maps[m].road_tris, np.float32).reshape(-1, 8)[:, :6].reshape(-1, 2)
            if tris.shape[0] == 0:      # a map without a road mesh: look at the ego
                camera_xy = tuple(float(v) for v in self.get_state()[int(env), 0, :2])
            else:
                camera_xy = (0.5 * (float(tris[:, 0].min()) + float(tris[:, 0].max())), 0.5 * (float(tris[:, 1].min()) + float(tris[:, 1].max())))
        if out is None:
            out = torch.empty((3, H, W), dtype=torch.uint8, device=self.device)
        if out.shape != (3, H, W) or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != self.device:
            raise ValueError("render_view: out must be a contiguous uint8[3, H, W] tensor on the engine's device")
        self._check(self.lib.tde_render_view(self.h, int(env), float(camera_xy[0]), float(camera_xy[1]), float(camera_psi),
                                             float(fov), W, H, _ptr(out), self._stream()), "tde_render_view")
        return out

    def compute_infractions(self) -> torch.Tensor:
        self._check(self.lib.tde_compute_infractions(self.h, self._stream()), "tde_compute_infractions")
        return self.get_infractions()

    # -- state access
    def _get(self, fn, shape, dtype, what):
        out = torch.empty(shape, dtype=dtype, device=self.device)
        self._check(fn(self.h, _ptr(out), self._stream()), what)
        return out

    def _set(self, fn, t, shape, dtype, what):
        t = torch.as_tensor(t).to(device=self.device, dtype=dtype).contiguous().view(shape)
        self._check(fn(self.h, _ptr(t), self._stream()), what)
        torch.cuda.current_stream(self.device).synchronize()  # `t` may be a temporary

    def get_state(self): return self._get(self.lib.tde_get_state, (self.E, self.A, 4), torch.float32, "tde_get_state")
    def set_state(self, t): self._set(self.lib.tde_set_state, t, (self.E, self.A, 4), torch.float32, "tde_set_state")
    def get_attributes(self): return self._get(self.lib.tde_get_attributes, (self.E, self.A, 4), torch.float32, "tde_get_attributes")
    def set_attributes(self, t): self._set(self.lib.tde_set_attributes, t, (self.E, self.A, 4), torch.float32, "tde_set_attributes")
    def get_infractions(self): return self._get(self.lib.tde_get_infractions, (self.E, self.A, 4), torch.float32, "tde_get_infractions")
    def get_env_vars(self): return self._get(self.lib.tde_get_env_vars, (self.E, 8), torch.int32, "tde_get_env_vars")
    def set_env_vars(self, t): self._set(self.lib.tde_set_env_vars, t, (self.E, 8), torch.int32, "tde_set_env_vars")

    def episode_stats(self, reset: bool = False) -> np.ndarray:
        out = (C.c_double * TDE_NUM_STATS)()
        self._check(self.lib.tde_get_episode_stats(self.h, out, int(reset), self._stream()), "tde_get_episode_stats")
        return np.asarray(list(out), np.float64)

    def num_kernel_launches(self) -> int:
        n = C.c_int64()
        self._check(self.lib.tde_num_kernel_launches(self.h, C.byref(n)), "tde_num_kernel_launches")
        return int(n.value)

    def sm_count(self) -> int:
        n = C.c_int32()
        self._check(self.lib.tde_device_sm_count(self.h, C.byref(n)), "tde_device_sm_count")
        return int(n.value)

    def map_info(self, map_id: int = 0) -> Dict[str, int]:
        out = (C.c_int32 * 8)()
        self._check(self.lib.tde_get_map_info(self.h, int(map_id), out), "tde_get_map_info")
        keys = ["road_tris", "mark_tris", "stoplines", "grid_nx", "grid_ny", "grid_items", "safe_cells", "render_prims"]
        return dict(zip(keys, [int(v) for v in out]))

    # -- stateless kernels (config C4)
    def collision_boxes(self, state: torch.Tensor, attr: torch.Tensor) -> torch.Tensor:
        st = state.to(device=self.device, dtype=torch.float32).contiguous()
        at = attr.to(device=self.device, dtype=torch.float32).contiguous()
        E, A = st.shape[:2]
        out = torch.empty((E, A), dtype=torch.float32, device=self.device)
        check(self.lib, None, self.lib.tde_collision_boxes(_ptr(st), _ptr(at), E, A, _ptr(out), self._stream()), "tde_collision_boxes")
        return out

    def offroad_boxes(self, map_id: int, state: torch.Tensor, attr: torch.Tensor) -> torch.Tensor:
        st = state.to(device=self.device, dtype=torch.float32).contiguous()
        at = attr.to(device=self.device, dtype=torch.float32).contiguous()
        E, A = st.shape[:2]
        out = torch.empty((E, A), dtype=torch.float32, device=self.device)
        self._check(self.lib.tde_offroad_boxes(self.h, int(map_id), _ptr(st), _ptr(at), E, A, _ptr(out), self._stream()), "tde_offroad_boxes")
        return out


def info_dict(info_row) -> Dict[str, float]:
    return {k: float(info_row[i]) for k, i in INFO_COLUMNS.items()}
