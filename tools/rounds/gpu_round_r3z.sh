#!/bin/bash
# Round 2, call AZ: ncu --set full of the final C4 kernels (one launch each) and the C4 / C2 bench lines of the final build.
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tde_\(collision\|offroad\)_kernel -s 6 -c 2 -o gpurun_out/prof_c4 -f \
    python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_c4.log 2>&1
python bench.py --workload c4 --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4.json 2>/dev/null
python bench.py --workload c2 --steps 3000 --warmup 100 --no-cpu-baseline > gpurun_out/bench_c2.json 2>/dev/null
tail -c 300 gpurun_out/bench_c4.json
