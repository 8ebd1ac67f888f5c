#!/bin/bash
# Round 2, call E: register budget / occupancy variants of both kernels; ncu of the C4 micro-benchmark kernels.
set -x
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
export TDE_PHYS_STAGE=0
for lib in base nw1 nw1r80 nw1r96 nw2r80p80 p96; do
  if [ "$lib" = base ]; then unset TDE_B200_LIB; else export TDE_B200_LIB=$PWD/variants/lib_$lib.so; fi
  echo "== $lib" | tee -a gpurun_out/ab.txt
  python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
done
unset TDE_B200_LIB
ncu --set full --clock-control none --import-source on -k regex:tde_\(collision\|offroad\)_kernel -s 6 -c 2 -o gpurun_out/prof_c4 -f \
    python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_c4.log 2>&1
python bench.py --workload c4 --steps 50 --warmup 5 --no-cpu-baseline | tee gpurun_out/bench_c4.json
python bench.py --workload c2 --steps 200 --warmup 20 --no-cpu-baseline | tee gpurun_out/bench_c2.json
ls -la gpurun_out
