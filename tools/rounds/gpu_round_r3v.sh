#!/bin/bash
# Round 2, call AV: VecFrameStack as a ring of stacked observations (tde_step_stacked_ring) against the in-place shift.
set -x
timeout 900 python -m pytest tests/test_gpu_env_api.py tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
python tools/kernel_times.py | head -1
python tools/kernel_times.py 8192 8 | head -1
python tools/kernel_times.py 1024 16 | head -1
