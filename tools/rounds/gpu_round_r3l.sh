#!/bin/bash
# Round 2, call AL: per-env trace of the physics kernel with fixed stride / tickets after the env / tickets before the env.
set -x
for v in trace tick1trace tick2trace; do
  echo "== $v 16384"; TDE_B200_LIB=$PWD/variants/lib_$v.so python tools/trace_envs.py 2>&1 | grep -A13 "^physics"
  echo "== $v 1024"; TDE_B200_LIB=$PWD/variants/lib_$v.so python tools/trace_envs.py 1024 16 2>&1 | grep -A13 "^physics"
done
for v in tick1 tick2; do echo "== $v"; TDE_B200_LIB=$PWD/variants/lib_$v.so python tools/kernel_times.py | head -1; TDE_B200_LIB=$PWD/variants/lib_$v.so python tools/kernel_times.py 1024 16 | head -1; done
