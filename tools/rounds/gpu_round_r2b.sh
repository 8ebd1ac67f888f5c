#!/bin/bash
# Round 2, call B: staged physics (bulk-async copies) on/off, render groups of 1 / 2 / 4 warps with the barrier fix.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
tools/ab_checked.sh base nw1 nw4
echo "== base, TDE_PHYS_STAGE=0" | tee -a gpurun_out/ab.txt
TDE_PHYS_STAGE=0 python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
echo "== base, TDE_PHYS_WARPS=16" | tee -a gpurun_out/ab.txt
TDE_PHYS_WARPS=16 python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
echo "== base, TDE_PHYS_WARPS=8" | tee -a gpurun_out/ab.txt
TDE_PHYS_WARPS=8 python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
TDE_B200_LIB=$PWD/variants/lib_trace.so python tools/trace_envs.py 2>&1 | tee gpurun_out/trace.txt
python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/bench_err.txt | tee gpurun_out/bench.json
ncu --set full --clock-control none --import-source on -k regex:tde_.*_kernel -s 13 -c 2 -o gpurun_out/prof_step -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_bench2.log 2>&1
ls -la gpurun_out
