"""Shared by the CPU and GPU tests of the reference-made fixtures (tests/golden/ref_*.npz, generated
by tests/golden/make_reference_golden.py from the reference's own gym_env.py)."""
import glob
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "ref_*.npz")) + glob.glob(os.path.join(HERE, "golden", "refsuite_*.npz")))

# the reference evaluates reward and info in float64 (math.dist / math.cos on fp32 tensors, gym_env.py:401-403,
# :430-435); the path computes them in binary32 (DESIGN D12): north-star tolerance 1e-5 relative
RTOL = 1e-5


def load(name):
    d = dict(np.load(os.path.join(HERE, "golden", name + ".npz")))
    d["env_config"] = json.loads(str(d["env_config"]))
    d["info_keys"] = json.loads(str(d["info_keys"]))
    d["types"] = json.loads(str(d["types"]))
    return d


def scenario_set(d):
    from torchdriveenv_b200 import scenarios as S
    if "suite_waypoints" in d:
        # an entry of the reference's validation_cases.yml (frozen in the fixture): WaypointSuite -> scenario tables through
        # the product's own scenario_set_from_suite, as WaypointSuiteEnv does
        from torchdriveenv_b200 import gym_env as G
        sc = None
        if len(d["suite_agent_states"]):
            sc = G.Scenario(agent_states=d["suite_agent_states"].tolist(), agent_attributes=d["suite_agent_attributes"].tolist(),
                            recurrent_states=[[0.0]] * len(d["suite_agent_states"]))
        seqs = {int(q): d["suite_car_seqs"][i].tolist() for i, q in enumerate(d["suite_car_seq_keys"])}
        suite = G.WaypointSuite(locations=[str(d["suite_location"])], waypoint_suite=[d["suite_waypoints"].tolist()],
                                car_sequence_suite=[seqs], scenarios=[sc])
        ss = G.scenario_set_from_suite(G.EnvConfig(**d["env_config"]), suite, n_background=0, seed=0)
        return ss, ss.max_agents()
    builders = {"three_way": lambda: S.three_way(6), "traffic_lights": lambda: S.traffic_lights(12),
                "roundabout": lambda: S.roundabout(8), "validation_mix": lambda: S.validation_mix(8)}
    return builders[str(d["scenario_set"])](), int(d["max_agents"])


def engine_kwargs(d):
    """The fixture's EnvConfig overrides -> tde_config fields, through the product's own mapping."""
    from torchdriveenv_b200 import gym_env as G
    return G.engine_config(G.EnvConfig(**d["env_config"]), auto_reset=0)


def check_against_reference(d, reward, terminated, truncated, info, target_idx, states):
    """reward[T], terminated[T], truncated[T], info[T][>=9], target_idx[T], states[T][4] of the path under test
    against what the reference's WaypointSuiteEnv returned for the same episode."""
    T = len(d["actions"])
    np.testing.assert_allclose(states, d["states"], rtol=RTOL, atol=0, err_msg="ego state")
    assert np.array_equal(np.asarray(terminated, np.uint8), d["terminated"]), "terminated (is_terminated, gym_env.py:413-417)"
    assert np.array_equal(np.asarray(truncated, np.uint8), d["truncated"]), "truncated (is_truncated, gym_env.py:134-135)"
    scale = np.maximum(1.0, np.abs(d["reward"]))
    assert np.max(np.abs(np.asarray(reward, np.float64) - d["reward"]) / scale) <= RTOL, "reward (get_reward, gym_env.py:396-411)"
    want = d["info"]
    got = np.asarray(info, np.float64)[:, :want.shape[1]]
    for k, key in enumerate(d["info_keys"]):
        sc = np.maximum(1.0, np.abs(want[:, k]))
        assert np.max(np.abs(got[:, k] - want[:, k]) / sc) <= RTOL, f"info[{key}] (get_info, gym_env.py:419-437)"
    assert np.array_equal(np.asarray(target_idx, np.int32), d["target_idx"]), "current_target_idx (gym_env.py:378-383)"
    assert T == len(reward)
