#!/bin/bash
# Round 2, call D: GPU tests with tde_clone / device guard; step split over two internal streams (tail overlap).
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
rm -f gpurun_out/ab.txt
for lib in base nw1; do
  if [ "$lib" = base ]; then unset TDE_B200_LIB; else export TDE_B200_LIB=$PWD/variants/lib_$lib.so; fi
  for v in "TDE_PHYS_STAGE=0 TDE_STEP_SPLIT=1" "TDE_PHYS_STAGE=0 TDE_STEP_SPLIT=2" "TDE_PHYS_STAGE=0 TDE_STEP_SPLIT=4" "TDE_PHYS_STAGE=0 TDE_STEP_SPLIT=8" "TDE_PHYS_STAGE=0 TDE_STEP_SPLIT=16" "TDE_PHYS_STAGE=1 TDE_STEP_SPLIT=4"; do
    echo "== $lib, $v" | tee -a gpurun_out/ab.txt
    env $v python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
  done
done
ls -la gpurun_out
