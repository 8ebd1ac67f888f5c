#!/bin/bash
# Round 2, call U: static candidates bounded per tile row by the turned viewport (base) against the bounding square (prev).
set -x
tools/ab_checked.sh prev base prev base
TDE_B200_LIB=$PWD/variants/lib_prev.so python tools/kernel_times.py 8192 8 | head -1
python tools/kernel_times.py 8192 8 | head -1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_env_api.py -x -q 2>&1 | tail -2
