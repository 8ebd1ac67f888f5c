#!/bin/bash
# Round 2, call R: per-env trace of one C3 step (env durations, envs in flight over time) with the current build.
set -x
mkdir -p gpurun_out
TDE_B200_LIB=variants/lib_trace.so python tools/trace_envs.py 2>&1 | tail -40
