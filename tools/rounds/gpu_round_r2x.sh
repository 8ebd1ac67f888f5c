#!/bin/bash
# Round 2, call X: cheaper SAT broad phase (FMA, pre-scaled radii, pair mask after the loop), collision grid from the
# occupancy calculator, flattened offroad; against the previous build.
set -x
tools/ab_checked.sh prev base prev base
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q 2>&1 | tail -2
python tools/c4_times.py
TDE_B200_LIB=$PWD/variants/lib_prev.so python tools/c4_times.py
python tools/kernel_times.py 1024 16 | head -1
TDE_B200_LIB=$PWD/variants/lib_prev.so python tools/kernel_times.py 1024 16 | head -1
ncu --set full --clock-control none --import-source on -k 'regex:tde_offroad_kernel' -s 4 -c 1 -o gpurun_out/prof_c4_offroad -f \
    python tools/c4_times.py > gpurun_out/ncu_c4.log 2>&1
