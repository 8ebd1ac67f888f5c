#!/bin/bash
# ncu full capture of one physics + one render launch of the C3 bench
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:tde_.*_kernel' -s 13 -c 2 -o gpurun_out/prof_step -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_bench2.log 2>&1
tail -2 gpurun_out/ncu_bench2.log | cut -c1-300
