#!/bin/bash
# Round 2, call AH: static candidates from the bounding box of the turned viewport (base) against that of its circumcircle (prev).
set -x
tools/ab_checked.sh prev base prev base
python tools/kernel_times.py 8192 8 | head -1
TDE_B200_LIB=$PWD/variants/lib_prev.so python tools/kernel_times.py 8192 8 | head -1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_env_api.py -x -q 2>&1 | tail -2
