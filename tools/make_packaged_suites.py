"""Packages the reference's bundled waypoint suites (torchdriveenv/data/validation_cases.yml, training_cases.yml: scenario
INPUT DATA - locations, waypoint polylines, replay car sequences, predetermined agents) as compact JSON
under torchdriveenv_b200/data/, so that load_default_validation_data() / load_default_train_data() work without the
reference checkout.  Values are kept exactly as the YAML holds them.  Usage: python tools/make_packaged_suites.py"""
import json
import os
import sys

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/torchdriveenv/data"
for name in ("validation_cases", "training_cases"):
    with open(os.path.join(REF, name + ".yml")) as f:
        d = yaml.load(f, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))
    out = dict(source=f"torchdriveenv/data/{name}.yml of inverted-ai/torchdriveenv @ 7b484bc (scenario input data)",
               locations=d["locations"], waypoint_suite=d["waypoint_suite"], car_sequence_suite=d["car_sequence_suite"],
               scenarios=d["scenarios"])
    path = os.path.join(ROOT, "torchdriveenv_b200", "data", name + ".json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes,", len(d["locations"]), "cases")
