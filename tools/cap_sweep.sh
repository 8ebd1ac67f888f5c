#!/bin/bash
# resident-blocks-per-SM sweep of the two step kernels (knobs read by configure_kernels)
mkdir -p gpurun_out; rm -f gpurun_out/caps.txt
for pc in 8 7; do for rc in 8 7; do
  echo "== phys cap $pc render cap $rc" | tee -a gpurun_out/caps.txt
  TDE_PHYS_BLOCKS_CAP=$pc TDE_RENDER_BLOCKS_CAP=$rc python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/caps.txt
done; done
