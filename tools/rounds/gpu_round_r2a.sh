#!/bin/bash
# Round 2, call A: GPU tests on the in-tree library, parity-gated A/B of the render variants (round-1 library, 1 / 2 / 4
# warps per env), per-env trace, ncu launch list + full capture of one physics + one render launch, bench.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/smi.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
tools/ab_checked.sh r01 base nw1 nw4
TDE_B200_LIB=$PWD/variants/lib_trace.so python tools/trace_envs.py 2>&1 | tee gpurun_out/trace.txt
python bench.py --steps 200 --warmup 20 2> gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_bench1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tde_.*_kernel -s 13 -c 2 -o gpurun_out/prof_step -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_bench2.log 2>&1
ls -la gpurun_out
