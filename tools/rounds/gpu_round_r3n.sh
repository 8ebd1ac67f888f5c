#!/bin/bash
# Round 2, call AN: physics tickets drawn mid-env from one counter per group of CTAs (pt5) against the fixed stride.
set -x
tools/ab_checked.sh base pt5 base pt5
for v in base pt5; do
  if [ "$v" = base ]; then unset TDE_B200_LIB; else export TDE_B200_LIB=$PWD/variants/lib_$v.so; fi
  echo "== $v"; python tools/kernel_times.py 8192 8 | head -1; python tools/kernel_times.py 1024 16 | head -1; python tools/kernel_times.py 4096 32 | head -1
done
TDE_B200_LIB=$PWD/variants/lib_pt5trace.so python tools/trace_envs.py 2>&1 | grep -A13 "^physics"
TDE_B200_LIB=$PWD/variants/lib_pt5.so timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_env_api.py -x -q 2>&1 | tail -1
