#!/bin/bash
# Round 2, call BB: the gym.Env step of one env through tde_step_host (one synchronisation per step): tests and the C1 line.
set -x
timeout 900 python -m pytest tests/test_gpu_env_api.py tests/test_gpu_reference_golden.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
python - <<'PY'
import json, bench
print(json.dumps(bench.run_c1_gym_api(0)))
PY
