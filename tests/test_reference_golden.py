"""The oracle against fixtures made by the REFERENCE'S OWN CODE: tests/golden/ref_*.npz hold what the
unmodified `torchdriveenv/gym_env.py` (WaypointSuiteEnv + SingleAgentWrapper) returned for whole
episodes when driven over a SimulatorInterface-level surface (tests/golden/make_reference_golden.py).
This pins SURVEY §8 rows a1 (step ordering), a9 (reward), a10 (waypoint progress), a11 (termination,
truncation), a12 (info) and a13 (spaces, output conventions) of the oracle to the reference itself."""
import json
import os

import numpy as np
import pytest

import reference_golden_util as R


def test_fixtures_exist_and_cover_the_decisions():
    assert len(R.NAMES) >= 15 and sum(n.startswith("refsuite_") for n in R.NAMES) == 5
    seen = dict(terminated=0, truncated=0, reached=0, offroad=0, collision=0, red=0)
    for n in R.NAMES:
        d = R.load(n)
        seen["terminated"] += int(d["terminated"].sum()); seen["truncated"] += int(d["truncated"].sum())
        seen["reached"] += int(d["info"][-1, 4]); seen["offroad"] += int((d["info"][:, 0] > 0).sum())
        seen["collision"] += int((d["info"][:, 1] > 0).sum()); seen["red"] += int((d["info"][:, 2] > 0).sum())
    assert all(v > 0 for v in seen.values()), seen


@pytest.mark.parametrize("name", R.NAMES)
def test_oracle_matches_reference_episode(oracle, name):
    from torchdriveenv_b200._capi import default_config
    d = R.load(name)
    ss, A = R.scenario_set(d)
    orc = oracle.OracleEnvSet(default_config(num_envs=1, max_agents=A, **R.engine_kwargs(d)), ss.pack(A))
    k = int(d["scenario"])
    orc.set_env_scenario_range([k], [k + 1])
    orc.reset(seed=int(d["seed"]))
    orc.state[0, 0, :] = d["start_state"]          # the start pose the reference's reset() drew (set_start_pos :351-367)
    rew, term, trunc, info, tgt, states = [], [], [], [], [], []
    for a in d["actions"]:
        _, r, te, tr, inf = orc.step(a[None])
        rew.append(r[0]); term.append(te[0]); trunc.append(tr[0]); info.append(inf[0].copy())
        tgt.append(orc.env_vars[0, 2]); states.append(orc.state[0, 0].copy())
    assert np.array_equal(np.asarray(states), d["states"]), "oracle trajectory changed since the fixture was made"
    R.check_against_reference(d, rew, term, trunc, info, tgt, np.asarray(states))


def test_reference_output_conventions():
    """What SingleAgentWrapper.step hands back (gym_env.py:453-461): obs uint8[3,64,64], python float reward,
    python bools, an info dict of 0-d tensors / scalars with exactly the keys of get_info :426-436."""
    t = R.load(R.NAMES[0])["types"]
    assert t["obs"] == ["uint8", [3, 64, 64]]
    assert t["reward"] == "float" and t["terminated"] == "bool" and t["truncated"] == "bool"
    assert list(t["info"]) == R.load(R.NAMES[0])["info_keys"]
    assert t["info"]["offroad"] == "tensor[]" and t["info"]["reached_waypoint_num"] == "int"


@pytest.mark.skipif(not os.path.exists("/root/reference/torchdriveenv/gym_env.py"), reason="reference checkout not present")
def test_fixtures_regenerate_from_the_reference(oracle):
    """Where the reference checkout exists: re-run its code and compare with the committed vectors."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(R.HERE, "golden", "make_reference_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    ref = m.import_reference()
    for name in ("ref_three_way_swerve_offroad", "ref_mix_custom_rewards"):
        res, report = m.run_case(ref, name)
        d = R.load(name)
        for key in ("actions", "reward", "terminated", "truncated", "info", "states", "target_idx", "start_state"):
            assert np.array_equal(res[key], d[key]), f"{name}:{key}"
        assert report["flags_equal"] and report["reward_max_abs"] < 1e-5 and report["info_max_abs"] < 1e-5
    # the reference's own validation suite, loaded by the reference's loader (and by the product's: same content), with the
    # arguments the reference hands to build_simulator checked against the product's scenario tables
    ref_utils = m.import_reference_env_utils()
    for name in ("refsuite_case1_parked_car", "refsuite_case2_chicken"):
        res, report = m.run_suite_case(ref, ref_utils, name)
        d = R.load(name)
        for key in ("actions", "reward", "terminated", "truncated", "info", "states", "target_idx", "start_state", "suite_waypoints", "suite_car_seqs"):
            assert np.array_equal(res[key], d[key]), f"{name}:{key}"
        assert report["flags_equal"] and report["reward_max_abs"] < 1e-5 and report["info_max_abs"] < 1e-5


def _metrics_fixture():
    d = dict(np.load(os.path.join(R.HERE, "golden", "refmetrics_validation_mix.npz")))
    return d, json.loads(str(d["ref_metrics"]))


def test_episode_statistics_match_the_reference_callback(oracle):
    """The episode-statistics vector (TDE_STAT_*, what the N GPUs all-reduce) against the counters of the reference's
    own EvalNTimestepsCallback._calc_metrics (examples/rl_training.py:39-67), which were fed every finished episode
    of this run when the fixture was made (tests/golden/make_reference_golden.py::run_metrics_case)."""
    from torchdriveenv_b200 import scenarios as S
    from torchdriveenv_b200._capi import STAT_NAMES, default_config
    d, want = _metrics_fixture()
    E, A = int(d["num_envs"]), int(d["max_agents"])
    orc = oracle.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1, max_environment_steps=60), S.validation_mix(8).pack(A))
    orc.reset(seed=int(d["seed"]))
    for a in d["actions"]:
        orc.step(a)
    got = {n: float(orc.stats[i]) for i, n in enumerate(STAT_NAMES)}
    assert want["episodes"] > 300
    for k, v in want.items():
        assert got[k] == float(v), k


@pytest.mark.skipif(not os.path.exists("/root/reference/examples/rl_training.py"), reason="reference checkout not present")
def test_metrics_fixture_regenerates_from_the_reference(oracle):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(R.HERE, "golden", "make_reference_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    rl = m.import_reference_eval_callback()
    res, want, stats = m.run_metrics_case(rl)
    d, frozen = _metrics_fixture()
    assert want == frozen and np.array_equal(res["actions"], d["actions"])
    assert all(float(want[k]) == stats[k] for k in want)


def test_labeled_data_loader_matches_the_reference():
    """env_utils.load_labeled_data (scenario-builder JSON, reference env_utils.py:31-105): tests/golden/ref_labeled_suite.json
    is what the reference's own loader returned for tests/golden/labeled_json/*.json with random.seed(7)."""
    import random
    from torchdriveenv_b200 import env_utils as U
    want = json.load(open(os.path.join(R.HERE, "golden", "ref_labeled_suite.json")))
    random.seed(7)
    suite = U.load_labeled_data(os.path.join(R.HERE, "golden", "labeled_json"))
    rows = []
    for k, loc in enumerate(suite.locations):
        sc, seqs = suite.scenarios[k], suite.car_sequence_suite[k]
        rows.append(dict(location=loc, waypoints=suite.waypoint_suite[k],
                         agent_states=None if sc is None else sc.agent_states, agent_attributes=None if sc is None else sc.agent_attributes,
                         n_recurrent=None if sc is None else [len(r) for r in sc.recurrent_states],
                         car_sequences=None if seqs is None else {str(q): v for q, v in sorted(seqs.items())}))
    got = json.loads(json.dumps(sorted(rows, key=lambda r: r["location"])))
    assert got == want
    assert want[1]["car_sequences"].keys() == {"1", "2"} and len(want[1]["car_sequences"]["1"]) == 200 and len(want[1]["car_sequences"]["2"]) == 5
    assert 5 <= want[2]["agent_states"][0][3] <= 10


def test_env_config_construction_matches_the_reference():
    """construct_env_config (reference env_utils.py:10-12) on the `env:` sections of the reference's eight shipped training
    configs: tests/golden/ref_env_configs.json holds the raw sections and every EnvConfig field the reference ended up with."""
    from torchdriveenv_b200 import env_utils as U
    frozen = json.load(open(os.path.join(R.HERE, "golden", "ref_env_configs.json")))
    assert len(frozen) == 8
    for name, entry in frozen.items():
        cfg = U.construct_env_config(entry["raw"])
        for k, v in entry["fields"].items():
            assert getattr(cfg, k) == v, f"{name}: {k}"
