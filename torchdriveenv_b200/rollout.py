"""GPU-resident on-policy rollout collection over the lockstep step (BASELINE config C5).

The reference collects rollouts with stable-baselines3 (examples/rl_training.py:159-160,178-181,200:
``SubprocVecEnv`` + ``VecFrameStack(n_stack, channels_order="first")`` feeding ``PPO.learn`` ->
``collect_rollouts`` -> ``RolloutBuffer.add``): every step the stacked observation, action, reward and
episode-start flag of every env are copied into the buffer on the host.  Here the buffer lives in HBM
and the render kernel stores every new frame straight into the stacked observations it belongs to
(``tde_step_rollout_scatter``: the newest channel group of slot t + 1 and one group further down in each
of the next n_stack - 1 slots), so the buffer is filled without reading or moving a frame; reward and
flags are written by the physics kernel into the buffer's own rows.  ``frame_copy="shift"`` selects the
older ``tde_step_rollout`` (older frames read from slot t, shifted stack written to slot t + 1).
``frame_copy="ring"`` stores every frame exactly once: the buffer keeps single frames ([T + n_stack, E, 3, 64, 64],
observation t = frames t .. t + n_stack - 1) and an age per observation (how many of the older frames belong to the
running episode); ``RolloutBuffer.stacked(t, envs)`` gathers the VecFrameStack observation where it is consumed, so a
rollout step writes 12 KB per env instead of 36.

Field names and shapes follow SB3's ``RolloutBuffer`` ([n_steps, n_envs, ...]); ``episode_starts[t]``
is 1 where ``observations[t]`` is the first observation of an episode.  ``returns_and_advantages`` is
GAE(lambda) as ``RolloutBuffer.compute_returns_and_advantage`` defines it.  stable-baselines3 itself
is not needed (and not installed in the build image).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch

from ._capi import TDE_INFO_STRIDE, TDE_OBS_H, TDE_OBS_W
from .engine import Engine


class RolloutBuffer:
    """[n_steps (+1 for observations / episode_starts), E, ...] tensors on the engine's GPU."""

    def __init__(self, n_steps: int, num_envs: int, n_stack: int, device, with_info: bool = False, ring: bool = False):
        T, E = int(n_steps), int(num_envs)
        self.n_steps, self.num_envs, self.n_stack = T, E, int(n_stack)
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
        if ring:
            # single frames: observation t is frames[t : t + n_stack] (oldest first); ages[t, e] = how many of the older
            # frames of observation t belong to env e's running episode (the others read as zeros, as VecFrameStack has them)
            self.observations = None
            self.frames = z((T + self.n_stack, E, 3, TDE_OBS_H, TDE_OBS_W), torch.uint8)
            self.ages = z((T + 1, E), torch.uint8)
        else:
            self.observations = z((T + 1, E, 3 * self.n_stack, TDE_OBS_H, TDE_OBS_W), torch.uint8)
            self.frames = self.ages = None
        self.actions = z((T, E, 2), torch.float32)
        self.rewards = z((T, E), torch.float32)
        self.terminated = z((T, E), torch.uint8)
        self.truncated = z((T, E), torch.uint8)
        self.episode_starts = z((T + 1, E), torch.uint8)
        self.values = z((T, E), torch.float32)
        self.log_probs = z((T, E), torch.float32)
        # info rows (get_info :419-437) per step are optional: 64 B per env-step
        self.infos = z((T, E, TDE_INFO_STRIDE), torch.float32) if with_info else None

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in vars(self).values() if isinstance(t, torch.Tensor))

    def stacked(self, t: int, envs: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Observation t as VecFrameStack hands it out, uint8[E or len(envs), 3 * n_stack, 64, 64]: a view of the stored
        stack, or (ring of frames) gathered from the n_stack frames with the ones from before the episode zeroed."""
        if self.observations is not None:
            o = self.observations[t]
            return o if envs is None else o[envs]
        n = self.n_stack
        age = self.ages[t] if envs is None else self.ages[t][envs]
        rows = age.shape[0]
        if out is None:
            out = torch.empty((rows, 3 * n, TDE_OBS_H, TDE_OBS_W), dtype=torch.uint8, device=self.frames.device)
        groups = out.view(rows, n, 3, TDE_OBS_H, TDE_OBS_W)
        for j in range(n):
            f = self.frames[t + j] if envs is None else self.frames[t + j][envs]
            keep = (age >= n - 1 - j).to(torch.uint8).view(rows, 1, 1, 1)      # frame j is n - 1 - j steps old
            torch.mul(f, keep, out=groups[:, j])
        return out


Policy = Callable[[torch.Tensor], object]


class RolloutCollector:
    """``collect(policy)`` fills a RolloutBuffer with n_steps lockstep steps of all E envs.

    ``policy(obs uint8[E, 3*n_stack, 64, 64]) -> actions[E, 2]`` or ``(actions, values[E], log_probs[E])``,
    all on the engine's GPU.  Finished envs restart inside the step kernel (auto-reset) and their frame
    stack restarts with zeros, as VecFrameStack does; the collector carries the last observation and
    episode-start flags over to slot 0 of the next rollout.
    """

    def __init__(self, engine: Engine, n_steps: int, n_stack: int = 3, seed: int = 0, with_info: bool = False,
                 frame_copy: str = "scatter", cuda_graph: bool = False):
        if n_stack < 2 or n_stack > 8:
            raise ValueError("n_stack must be in 2..8 (use Engine.step for a plain observation)")
        if frame_copy not in ("scatter", "shift", "ring"):
            raise ValueError("frame_copy must be 'scatter', 'shift' or 'ring'")
        self.frame_copy = frame_copy
        # cuda_graph=True: from the third rollout on, collect() replays one captured CUDA graph of the whole rollout (the
        # tde_* calls only enqueue work on the current stream, so they capture).  For policies made of static-shape GPU ops
        # that read their inputs from fixed tensors; a policy's `before_rollout(n_steps)` hook runs outside the graph (the
        # place for anything that cannot be captured, e.g. drawing random numbers from a torch.Generator).
        self.cuda_graph = bool(cuda_graph)
        self._graph = None
        self._graph_policy = None
        self.engine, self.n_steps, self.n_stack = engine, int(n_steps), int(n_stack)
        self.buffer = RolloutBuffer(n_steps, engine.E, n_stack, engine.device, with_info=with_info, ring=frame_copy == "ring")
        if frame_copy == "ring":
            self._slots = torch.arange(self.n_steps + 1, dtype=torch.int32, device=engine.device).view(-1, 1)
            self._stack_now = None      # scratch for policies that read the observation while the rollout is collected
        self._info = None if with_info else torch.zeros((engine.E, TDE_INFO_STRIDE), dtype=torch.float32, device=engine.device)
        self._seed = int(seed)
        self._started = False
        self.num_timesteps = 0

    def reset(self) -> torch.Tensor:
        b = self.buffer
        self.engine.reset(seed=self._seed)
        if self.frame_copy == "ring":
            b.frames[: self.n_stack].zero_()
            self.engine.render(out=b.frames[self.n_stack - 1])
            b.episode_starts[0].fill_(1)
            b.ages[0].zero_()
            self._started = True
            return b.stacked(0)
        b.observations[0].zero_()
        self.engine.render_stacked(b.observations[0], self.n_stack)
        b.episode_starts[0].fill_(1)
        self._started = True
        self._seed_older_groups()
        return b.observations[0]

    def _seed_older_groups(self) -> None:
        """Scatter mode: the frames of slot 0 also belong to the next n_stack - 1 stacked observations (one channel group
        further down each); once per rollout they are put there, after that the render kernel keeps the slots complete."""
        if self.frame_copy != "scatter":
            return
        b, n = self.buffer, self.n_stack
        for j in range(1, min(n, self.n_steps + 1)):
            b.observations[j][:, : 3 * (n - j)].copy_(b.observations[0][:, 3 * j:])

    def collect(self, policy: Policy) -> RolloutBuffer:
        if not self._started:
            self.reset()
        hook = getattr(policy, "before_rollout", None)
        if hook is not None:
            hook(self.n_steps)
        if not (self.cuda_graph and self.num_timesteps):
            self._collect_body(policy, carry=self.num_timesteps > 0)      # first rollout (no carry-over) / plain mode
        else:
            if self._graph is None or self._graph_policy is not policy:
                torch.cuda.synchronize(self.engine.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._collect_body(policy, carry=True)
                self._graph, self._graph_policy = g, policy              # the capture does not run the work: replay it now
            self._graph.replay()
        self.num_timesteps += self.n_steps * self.engine.E
        return self.buffer

    def _collect_ring(self, policy: Policy, carry: bool) -> None:
        """One rollout into the ring of frames: step t stores its frame once, in slot t + n_stack."""
        b, eng, T, n = self.buffer, self.engine, self.n_steps, self.n_stack
        if carry:
            for j in range(n):                               # n_stack frames per rollout: 1 / n_steps of the traffic
                b.frames[j].copy_(b.frames[T + j])           # (ascending: a source slot is read before it is a destination)
            b.episode_starts[0].copy_(b.episode_starts[T])
            b.ages[0].copy_(b.ages[T])
        into = bool(getattr(policy, "writes_into", False))
        reads = bool(getattr(policy, "needs_observation", True))
        for t in range(T):
            if reads:
                if self._stack_now is None:
                    self._stack_now = torch.empty((eng.E, 3 * n, TDE_OBS_H, TDE_OBS_W), dtype=torch.uint8, device=b.frames.device)
                obs = b.stacked(t, out=self._stack_now)
            else:
                obs = b.frames[t + n - 1]                    # the policy only looks at shape and device
            if into:
                policy(obs, out=b.actions[t])
            else:
                out = policy(obs)
                if isinstance(out, (tuple, list)):
                    act, val, logp = out
                    b.values[t].copy_(val.reshape(-1))
                    b.log_probs[t].copy_(logp.reshape(-1))
                else:
                    act = out
                b.actions[t].copy_(act.reshape(eng.E, 2))
            eng.step_into(b.actions[t], b.frames[t + n], reward=b.rewards[t], terminated=b.terminated[t], truncated=b.truncated[t],
                          info=b.infos[t] if b.infos is not None else self._info)
            if reads:   # the next observation's age is needed before the rollout ends
                running = 1 - torch.bitwise_or(b.terminated[t], b.truncated[t])
                torch.mul(torch.clamp(b.ages[t] + 1, max=n - 1), running, out=b.ages[t + 1])
        torch.bitwise_or(b.terminated, b.truncated, out=b.episode_starts[1:])
        if not reads:
            # all ages of the rollout at once: age[t] = min(n - 1, t - slot of the last episode start at or before t), the
            # episode running in slot 0 having started ages[0] slots before it
            first = torch.where(b.episode_starts[:1].bool(), 0, -b.ages[:1].to(torch.int32))
            begun = torch.where(b.episode_starts[1:].bool(), self._slots[1:], -(1 << 20))
            last = torch.cummax(torch.cat((first, begun), 0), dim=0).values
            b.ages.copy_(torch.clamp(self._slots - last, max=n - 1))

    def _collect_body(self, policy: Policy, carry: bool) -> None:
        if self.frame_copy == "ring":
            return self._collect_ring(policy, carry)
        b, eng, T = self.buffer, self.engine, self.n_steps
        if carry:
            b.observations[0].copy_(b.observations[T])      # one slot per rollout: 1/n_steps of the traffic
            b.episode_starts[0].copy_(b.episode_starts[T])
            self._seed_older_groups()
        into = bool(getattr(policy, "writes_into", False))   # the policy stores its actions in the buffer row itself
        for t in range(T):
            if into:
                policy(b.observations[t], out=b.actions[t])
            else:
                out = policy(b.observations[t])
                if isinstance(out, (tuple, list)):
                    act, val, logp = out
                    b.values[t].copy_(val.reshape(-1))
                    b.log_probs[t].copy_(logp.reshape(-1))
                else:
                    act = out
                b.actions[t].copy_(act.reshape(eng.E, 2))
            if self.frame_copy == "scatter":
                eng.step_rollout_scatter(b.actions[t], b.observations, t, self.n_stack,
                                         reward=b.rewards[t], terminated=b.terminated[t], truncated=b.truncated[t],
                                         info=b.infos[t] if b.infos is not None else self._info)
            else:
                eng.step_rollout(b.actions[t], b.observations[t], b.observations[t + 1], self.n_stack,
                                 reward=b.rewards[t], terminated=b.terminated[t], truncated=b.truncated[t],
                                 info=b.infos[t] if b.infos is not None else self._info)
        torch.bitwise_or(b.terminated, b.truncated, out=b.episode_starts[1:])     # one launch per rollout, not per step

    def last_observation(self) -> torch.Tensor:
        return self.buffer.stacked(self.n_steps)

    def returns_and_advantages(self, last_values: torch.Tensor, gamma: float = 0.99, gae_lambda: float = 0.95) -> Tuple[torch.Tensor, torch.Tensor]:
        """GAE(lambda) over the collected rollout (SB3 RolloutBuffer.compute_returns_and_advantage):
        delta_t = r_t + gamma V_{t+1} (1 - start_{t+1}) - V_t;  A_t = delta_t + gamma lambda (1 - start_{t+1}) A_{t+1}."""
        b, T = self.buffer, self.n_steps
        adv = torch.zeros_like(b.rewards)
        last = torch.zeros(b.num_envs, dtype=torch.float32, device=b.rewards.device)
        nxt = last_values.reshape(-1).to(torch.float32)
        for t in reversed(range(T)):
            nonterminal = 1.0 - b.episode_starts[t + 1].to(torch.float32)
            delta = b.rewards[t] + gamma * nxt * nonterminal - b.values[t]
            last = delta + gamma * gae_lambda * nonterminal * last
            adv[t] = last
            nxt = b.values[t]
        return adv + b.values, adv

    def episode_statistics(self, reset: bool = False) -> Dict[str, float]:
        from ._capi import STAT_NAMES
        s = self.engine.episode_stats(reset=reset)
        return {n: float(s[i]) for i, n in enumerate(STAT_NAMES)}


def uniform_policy(action_low=(-1.0, -0.3), action_high=(1.0, 0.3), seed: int = 0) -> Policy:
    """U(low, high) actions drawn on the GPU (the reference action space, gym_env.py:83-84): the stand-in
    for the CnnPolicy of examples/rl_training.py when no trainer is attached.  The uniforms of a whole rollout are drawn in
    `before_rollout` (outside a CUDA graph) into one fixed tensor; a step scales its slice straight into the caller's row
    (`writes_into`): one small launch per step, capturable."""
    state: Dict[str, object] = {"gen": None, "u": None, "k": 0}

    def setup(device):
        state["gen"] = torch.Generator(device=device)
        state["gen"].manual_seed(seed)
        state["lo"] = torch.tensor(action_low, dtype=torch.float32, device=device)
        state["span"] = torch.tensor(action_high, dtype=torch.float32, device=device) - state["lo"]

    def before_rollout(n_steps: int, num_envs: Optional[int] = None, device=None) -> None:
        if state["u"] is not None:
            torch.rand(state["u"].shape, generator=state["gen"], device=state["u"].device, out=state["u"])
            state["k"] = 0
        state["want"] = int(n_steps)

    def policy(obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if state["gen"] is None:
            setup(obs.device)
        if state["u"] is None or state["k"] >= state["u"].shape[0]:
            n = max(int(state.get("want", 64)), 1)
            state["u"] = torch.rand((n, obs.shape[0], 2), generator=state["gen"], device=obs.device)
            state["k"] = 0
        k = state["k"]
        state["k"] += 1
        if out is None:
            return torch.addcmul(state["lo"], state["u"][k], state["span"])
        return torch.addcmul(state["lo"], state["u"][k], state["span"], out=out)

    policy.writes_into = True
    policy.needs_observation = False     # a ring-of-frames collector does not have to gather the stack for it
    policy.before_rollout = before_rollout
    return policy
