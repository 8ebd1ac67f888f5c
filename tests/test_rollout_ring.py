"""Host logic of the ring-of-frames rollout buffer (RolloutCollector(frame_copy="ring")) without a GPU: a stand-in engine
that emits numbered frames and scripted episode ends, against a direct VecFrameStack simulation
(reference: VecFrameStack(env, n_stack=3, channels_order="first"), examples/rl_training.py:160)."""
import numpy as np
import pytest
import torch

from torchdriveenv_b200._capi import TDE_INFO_STRIDE
from torchdriveenv_b200.rollout import RolloutCollector, uniform_policy


class ScriptedEngine:
    """Frames carry (env, frame counter) in their pixels; an env ends an episode when the script says so and the frame
    returned by that step is the first one of the next episode (auto-reset inside the step, as tde_step does it)."""

    def __init__(self, E, seed):
        self.E, self.device = E, torch.device("cpu")
        self.rng = np.random.default_rng(seed)
        self.count = np.zeros(E, np.int64)

    def _frame(self, out):
        for e in range(self.E):
            out[e].fill_(int((7 * e + 3 * self.count[e]) % 251) + 1)
            out[e, 0, 0, 0] = int(self.count[e] % 256)

    def reset(self, seed=0):
        self.count[:] = 0

    def render(self, out):
        self._frame(out)
        return out

    def step_into(self, actions, obs, reward=None, terminated=None, truncated=None, info=None):
        done = self.rng.random(self.E) < 0.3
        trunc = done & (self.rng.random(self.E) < 0.5)
        self.count += 1
        self.count[done] += 100          # a new episode: frames of a different family
        self._frame(obs)
        reward.copy_(torch.from_numpy(self.rng.random(self.E).astype(np.float32)))
        terminated.copy_(torch.from_numpy((done & ~trunc).astype(np.uint8)))
        truncated.copy_(torch.from_numpy(trunc.astype(np.uint8)))
        return obs, reward, terminated, truncated, info


def reading_policy(seen):
    def policy(obs):
        seen.append(obs.clone())
        return torch.zeros((obs.shape[0], 2))
    return policy


@pytest.mark.parametrize("n_stack,T", [(3, 7), (4, 2), (2, 5), (3, 1)])
@pytest.mark.parametrize("reads", [True, False])
def test_ring_of_frames_reproduces_vec_frame_stack(n_stack, T, reads):
    E = 9
    eng = ScriptedEngine(E, seed=n_stack * 10 + T)
    col = RolloutCollector(eng, T, n_stack=n_stack, frame_copy="ring")
    seen = []
    policy = reading_policy(seen) if reads else uniform_policy(seed=1)
    stack = torch.zeros((E, 3 * n_stack, 64, 64), dtype=torch.uint8)
    first = True
    for r in range(4):
        buf = col.collect(policy)
        if first:   # the reset observation: zeros and the first frame
            stack[:, -3:] = buf.frames[n_stack - 1]
            first = False
        for t in range(T + 1):
            assert torch.equal(buf.stacked(t), stack), f"rollout {r} observation {t}"
            sub = torch.tensor([E - 1, 0, 4])
            assert torch.equal(buf.stacked(t, envs=sub), stack[sub])
            if t == T:
                break
            done = (buf.terminated[t] | buf.truncated[t]).bool()
            assert torch.equal(buf.episode_starts[t + 1].bool(), done)
            stack = torch.cat((stack[:, 3:], buf.frames[t + n_stack]), 1)
            stack[done, : 3 * (n_stack - 1)] = 0          # VecFrameStack: a new episode starts from zeros + its first frame
        assert torch.equal(col.last_observation(), stack)
    if reads:       # what the policy saw while the rollouts were collected is what the buffer hands out afterwards
        assert len(seen) == 4 * T
        assert torch.equal(seen[-1], buf.stacked(T - 1))
