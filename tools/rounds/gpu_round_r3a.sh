#!/bin/bash
# Round 2, call AA: physics envs handed out by ticket (next ticket drawn before the current env) against the fixed stride.
set -x
tools/ab_checked.sh prev base prev base
python tools/kernel_times.py 8192 8 | head -1
TDE_B200_LIB=$PWD/variants/lib_prev.so python tools/kernel_times.py 8192 8 | head -1
python tools/kernel_times.py 1024 16 | head -1
TDE_B200_LIB=$PWD/variants/lib_prev.so python tools/kernel_times.py 1024 16 | head -1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_env_api.py -x -q 2>&1 | tail -2
