"""Scenario / config file loaders with the reference's names (torchdriveenv/env_utils.py:10-123).
OmegaConf is replaced by PyYAML (the files are plain YAML); the schemas are unchanged:
WaypointSuite YAML = locations / waypoint_suite / car_sequence_suite / scenarios
(validation_cases.yml:1,7,84,1290); scenario-builder JSON = individual_suggestions['0'].states and
predetermined_agents[id].{states, static_attributes} (env_utils.py:31-105)."""
from __future__ import annotations

import json
import os
import random
from typing import Optional

import yaml

from .gym_env import EnvConfig, Scenario, SimulatorConfig, WaypointSuite


def construct_env_config(raw_config) -> EnvConfig:
    raw = dict(raw_config)
    sim = raw.pop("simulator", None)
    cfg = EnvConfig(**raw)
    if isinstance(sim, dict):
        cfg.simulator = SimulatorConfig(**{k: v for k, v in sim.items() if k in SimulatorConfig.__dataclass_fields__})
    return cfg


def load_env_config(yaml_path) -> EnvConfig:
    with open(yaml_path) as f:
        return construct_env_config(yaml.safe_load(f))


def load_waypoint_suite_data(yaml_path) -> WaypointSuite:
    with open(yaml_path) as f:
        data = yaml.load(f, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))
    suite = WaypointSuite(**data)
    if suite.scenarios is not None:
        suite.scenarios = [Scenario(agent_states=s["agent_states"], agent_attributes=s["agent_attributes"],
                                    recurrent_states=s.get("recurrent_states")) if s is not None else None
                           for s in suite.scenarios]
    return suite


def load_labeled_data(data_dir) -> WaypointSuite:
    """Scenario-builder JSON directory -> WaypointSuite (env_utils.py:31-105)."""
    suite = WaypointSuite(locations=[], waypoint_suite=[], car_sequence_suite=[], scenarios=[])
    for json_file in sorted(os.listdir(data_dir)):
        if not json_file.endswith(".json"):
            continue
        suite.locations.append(json_file.split('_')[1])
        with open(os.path.join(data_dir, json_file)) as f:
            data = json.load(f)
        suite.waypoint_suite.append([[s['center']['x'], s['center']['y']] for s in data['individual_suggestions']['0']['states']])
        scenario, car_sequences = None, None
        agents = data.get("predetermined_agents")
        if agents is not None:
            states, attrs, rec = [], [], []
            for key in agents:
                ag = agents[key]
                speed = random.randint(5, 10) if len(ag['states']) == 1 else 0
                s0 = ag['states']['0']
                states.append([s0['center']['x'], s0['center']['y'], s0['orientation'], speed])
                sa = ag['static_attributes']
                attrs.append([sa['length'], sa['width'], sa['rear_axis_offset']])
                rec.append([0] * 132)
            if states:
                scenario = Scenario(agent_states=states, agent_attributes=attrs, recurrent_states=rec)
            car_sequences = {}
            for key in agents:
                ag = agents[key]
                s0 = ag['states']['0']
                if ag['static_attributes'].get("max_speed", None) == 0:
                    car_sequences[int(key)] = [[s0['center']['x'], s0['center']['y'], s0['orientation'], 0] for _ in range(200)]
                elif len(ag['states']) > 1:
                    car_sequences[int(key)] = [[ag['states'][i]['center']['x'], ag['states'][i]['center']['y'],
                                                ag['states'][i]['orientation'], 0] for i in ag['states']]
        suite.scenarios.append(scenario)
        suite.car_sequence_suite.append(car_sequences)
    return suite


def _suite_from_dict(data) -> WaypointSuite:
    suite = WaypointSuite(locations=data["locations"], waypoint_suite=data["waypoint_suite"],
                          car_sequence_suite=data["car_sequence_suite"], scenarios=data["scenarios"])
    if suite.car_sequence_suite is not None:   # JSON object keys are strings; the reference's YAML has int agent slots
        suite.car_sequence_suite = [None if c is None else {int(k): v for k, v in c.items()} for c in suite.car_sequence_suite]
    if suite.scenarios is not None:
        suite.scenarios = [Scenario(agent_states=s["agent_states"], agent_attributes=s["agent_attributes"],
                                    recurrent_states=s.get("recurrent_states")) if s is not None else None
                           for s in suite.scenarios]
    return suite


def _load_default_data(file_name) -> Optional[WaypointSuite]:
    """The suites the reference bundles (torchdriveenv/__init__.py:8, env_utils.py:108-123): a YAML of that name under
    $TORCHDRIVEENV_DATA wins; otherwise the packaged copy torchdriveenv_b200/data/<name>.json (the same content, made by
    tools/make_packaged_suites.py from the reference's data files)."""
    extra = os.environ.get("TORCHDRIVEENV_DATA")
    if extra and os.path.exists(os.path.join(extra, file_name)):
        return load_waypoint_suite_data(os.path.join(extra, file_name))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", os.path.splitext(file_name)[0] + ".json")
    if os.path.exists(path):
        with open(path) as f:
            return _suite_from_dict(json.load(f))
    return None


def load_default_validation_data():
    return _load_default_data("validation_cases.yml")


def load_default_train_data():
    return _load_default_data("training_cases.yml")


# ----------------------------------------------------------------------------- background traffic

def load_background_traffic(json_path) -> dict:
    """One background-traffic file of the reference (gym_env.py:200-214; files under
    torchdriveenv/resources/background_traffic named ``carla_<Town>_<density>_<seed>.json``).  Schema:
    ``location``, ``agent_density``, ``random_seed``, ``agent_states[{center{x,y}, orientation, speed}]``,
    ``agent_attributes[{length, width, rear_axis_offset}]``, ``recurrent_states`` (IAI-internal, ignored).
    Returns plain arrays: ``states`` (n, 4) x y psi v and ``attributes`` (n, 3) length width lr."""
    import numpy as np
    with open(json_path) as f:
        d = json.load(f)
    st = [[a["center"]["x"], a["center"]["y"], a["orientation"], a["speed"]] for a in d["agent_states"]]
    at = [[a["length"], a["width"], a["rear_axis_offset"]] for a in d["agent_attributes"]]
    if len(st) != len(at):
        raise ValueError(f"{json_path}: {len(st)} agent_states but {len(at)} agent_attributes")
    return dict(location=d.get("location"), agent_density=int(d.get("agent_density", 0)), random_seed=d.get("random_seed"),
                states=np.asarray(st, np.float32).reshape(-1, 4), attributes=np.asarray(at, np.float32).reshape(-1, 3))


def pick_background_traffic(directory, town: str, rng: Optional[random.Random] = None) -> Optional[dict]:
    """The reference's file choice (gym_env.py:203-214): a random file of this town whose agent count
    plus density stays under 100.  ``town`` is the part after ``carla_`` (e.g. ``Town03``)."""
    rng = rng or random
    names = sorted(n for n in os.listdir(directory) if n.endswith(".json") and len(n.split("_")) > 1 and n.split("_")[1] == town)
    rng.shuffle(names)
    for n in names:
        bt = load_background_traffic(os.path.join(directory, n))
        if len(bt["states"]) + bt["agent_density"] < 100:
            return bt
    return None


def background_agents_for_start(bt: dict, ego_xy, min_distance: float = 100.0, max_agents: Optional[int] = None):
    """The agents the reference keeps from a background-traffic file: those farther than 100 m from the
    ego start (gym_env.py:229-233; nearer ones are re-sampled by the Inverted AI service, which is not
    reachable offline).  Returns (states (k, 4), attributes (k, 3)); they are driven at constant
    velocity by the step kernel (no replay)."""
    import numpy as np
    d = np.hypot(bt["states"][:, 0] - float(ego_xy[0]), bt["states"][:, 1] - float(ego_xy[1]))
    keep = d > min_distance
    st, at = bt["states"][keep], bt["attributes"][keep]
    if max_agents is not None:
        st, at = st[:max_agents], at[:max_agents]
    return st, at
