"""Generates the known-answer fixtures under tests/golden/ from the CPU oracle.

The reference ships no golden vectors and cannot be imported here (torchdrivesim is absent), so these
freeze the ORACLE's outputs: they make regressions of the oracle visible and give the CUDA path a
fixed target that does not depend on the oracle library being rebuilt.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from torchdriveenv_b200 import scenarios as S  # noqa: E402
from torchdriveenv_b200._capi import default_config  # noqa: E402

CASES = {
    # name: (scenario builder, E, A, steps, config overrides)
    "three_way_e4": (lambda: S.three_way(6), 4, 9, 40, dict(auto_reset=0)),
    "traffic_lights_e8": (lambda: S.traffic_lights(32), 8, 32, 30, dict(auto_reset=1)),
    "roundabout_e8": (lambda: S.roundabout(16), 8, 16, 30, dict(auto_reset=1, randomize_ego_attributes=1)),
    "mix_e10_a44": (lambda: S.validation_mix(40), 10, 44, 24, dict(auto_reset=1, left_handed_coordinates=0)),
}


def actions_for(name, E, steps):
    rng = np.random.default_rng(abs(hash(name)) % (2**31) if False else sum(map(ord, name)))
    return np.stack([rng.uniform(-1, 1, (steps, E)), rng.uniform(-0.3, 0.3, (steps, E))], -1).astype(np.float32)


def run_case(name):
    builder, E, A, steps, over = CASES[name]
    ss = builder()
    cfg = default_config(num_envs=E, max_agents=A, **over)
    env = O.OracleEnvSet(cfg, ss.pack(A))
    env.reset(seed=1234)
    acts = actions_for(name, E, steps)
    out = dict(actions=acts, reset_state=env.state.copy(), reset_vars=env.env_vars.copy(), reset_obs=env.render())
    rew, term, trunc, info, states = [], [], [], [], []
    for k in range(steps):
        obs, r, te, tr, inf = env.step(acts[k])
        rew.append(r); term.append(te); trunc.append(tr); info.append(inf); states.append(env.state.copy())
    out.update(reward=np.stack(rew), terminated=np.stack(term), truncated=np.stack(trunc), info=np.stack(info),
               states=np.stack(states), final_obs=obs, final_infractions=env.infractions.copy(),
               final_vars=env.env_vars.copy(), stats=env.stats.copy())
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        data = run_case(name)
        np.savez_compressed(os.path.join(here, name + ".npz"), **data)
        print(name, {k: v.shape for k, v in data.items() if hasattr(v, "shape")})
