#!/bin/bash
# Round 2, call AB: ticket counters per group of CTAs (physics ticketed with a prefetched ticket, render) against the previous
# build (prev), physics back on the fixed stride (static), render drawing its next ticket at the start of the env (early).
set -x
tools/ab_checked.sh prev base static early base prev
python tools/kernel_times.py 8192 8 | head -1
TDE_B200_LIB=$PWD/variants/lib_prev.so python tools/kernel_times.py 8192 8 | head -1
python tools/kernel_times.py 1024 16 | head -1
TDE_B200_LIB=$PWD/variants/lib_prev.so python tools/kernel_times.py 1024 16 | head -1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_env_api.py -x -q 2>&1 | tail -2
