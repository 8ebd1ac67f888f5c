"""The CUDA path, through the C ABI, against fixtures made by the reference's own gym_env.py
(tests/golden/ref_*.npz, see tests/golden/make_reference_golden.py): reward, termination, truncation,
info and waypoint progress of whole episodes as the reference's WaypointSuiteEnv decided them."""
import numpy as np
import pytest
import torch

import reference_golden_util as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", R.NAMES)
def test_cuda_matches_reference_episode(name):
    from torchdriveenv_b200.engine import Engine
    d = R.load(name)
    ss, A = R.scenario_set(d)
    eng = Engine(ss, 1, A, device="cuda:0", **R.engine_kwargs(d))
    k = int(d["scenario"])
    eng.set_env_scenario_range(np.asarray([k], np.int32), np.asarray([k + 1], np.int32))
    eng.reset(seed=int(d["seed"]))
    st = eng.get_state()
    st[0, 0, :] = torch.from_numpy(d["start_state"]).to(st.device)    # the start pose the reference's reset() drew
    eng.set_state(st)
    rew, term, trunc, info, tgt, states = [], [], [], [], [], []
    for a in d["actions"]:
        _, r, te, tr, inf = eng.step(torch.from_numpy(a[None]).cuda())
        rew.append(float(r[0])); term.append(int(te[0])); trunc.append(int(tr[0])); info.append(inf[0].cpu().numpy().copy())
        tgt.append(int(eng.get_env_vars()[0, 2])); states.append(eng.get_state()[0, 0].cpu().numpy().copy())
    R.check_against_reference(d, rew, term, trunc, info, tgt, np.asarray(states))
    eng.close()


def test_product_api_returns_what_the_reference_returns():
    """SingleAgentWrapper(WaypointSuiteEnv).step of the product hands back the same python types, shapes and info
    keys as the reference's (recorded in the fixtures' `types`)."""
    from torchdriveenv_b200 import gym_env as G
    from torchdriveenv_b200 import scenarios as S
    d = R.load("ref_three_way_default")
    env = G.SingleAgentWrapper(G.WaypointSuiteEnv(G.EnvConfig(seed=3, device="cuda:0"), S.three_way(6)))
    obs, _ = env.reset()
    assert [str(obs.dtype), list(obs.shape)] == d["types"]["obs"]
    obs, reward, terminated, truncated, info = env.step(np.array([0.5, 0.0], np.float32))
    assert [str(obs.dtype), list(obs.shape)] == d["types"]["obs"]
    assert type(reward).__name__ == d["types"]["reward"]
    assert type(terminated).__name__ == d["types"]["terminated"] and type(truncated).__name__ == d["types"]["truncated"]
    for key, kind in d["types"]["info"].items():
        assert key in info, key
        got = f"tensor{list(info[key].shape)}" if torch.is_tensor(info[key]) else type(info[key]).__name__
        if key == "dist_reward":    # the reference returns the bonus (float) or the literal 0 (int), gym_env.py:434
            assert got in ("float", "int")
        else:
            assert got == kind, f"info[{key}]: {got} vs the reference's {kind}"
    env.close()


def test_cuda_episode_statistics_match_the_reference_callback():
    """tde_get_episode_stats against the counters of the reference's EvalNTimestepsCallback._calc_metrics
    (examples/rl_training.py:39-67) for the 366 episodes of the frozen run."""
    import json
    import os
    from torchdriveenv_b200 import scenarios as S
    from torchdriveenv_b200._capi import STAT_NAMES
    from torchdriveenv_b200.engine import Engine
    d = dict(np.load(os.path.join(R.HERE, "golden", "refmetrics_validation_mix.npz")))
    want = json.loads(str(d["ref_metrics"]))
    E, A = int(d["num_envs"]), int(d["max_agents"])
    eng = Engine(S.validation_mix(8), E, A, device="cuda:0", auto_reset=1, max_environment_steps=60)
    eng.reset(seed=int(d["seed"]))
    for a in d["actions"]:
        eng.step(torch.from_numpy(a).cuda(), render=False)
    got = {n: float(v) for n, v in zip(STAT_NAMES, eng.episode_stats())}
    for k, v in want.items():
        assert got[k] == float(v), k
    eng.close()
