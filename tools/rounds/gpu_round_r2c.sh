#!/bin/bash
# Round 2, call C: two-stage culling + prefetch; render 1 vs 2 warps per env; physics staged (ld.shared / generic loads) vs global.
set -x
mkdir -p gpurun_out
tools/ab_checked.sh base nw1 gld
for v in "TDE_PHYS_STAGE=0" "TDE_PHYS_STAGE=0 TDE_PHYS_CARVEOUT=50"; do
  echo "== base, $v" | tee -a gpurun_out/ab.txt
  env $v python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
done
TDE_B200_LIB=$PWD/variants/lib_trace.so python tools/trace_envs.py 2>&1 | tee gpurun_out/trace.txt
ncu --set full --clock-control none --import-source on -k regex:tde_.*_kernel -s 13 -c 2 -o gpurun_out/prof_step -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_bench2.log 2>&1
TDE_PHYS_STAGE=0 ncu --set full --clock-control none --import-source on -k regex:tde_physics_kernel -s 6 -c 1 -o gpurun_out/prof_phys_global -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_bench3.log 2>&1
ls -la gpurun_out
