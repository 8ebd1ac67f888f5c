#!/bin/bash
# Round 2, call W: ncu --set full of the C4 kernels (flattened offroad) with source correlation.
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:tde_(offroad|collision)_kernel' -s 6 -c 2 -o gpurun_out/prof_c4 -f \
    python tools/c4_times.py > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log | cut -c1-200
ls -la gpurun_out/prof_c4.ncu-rep
