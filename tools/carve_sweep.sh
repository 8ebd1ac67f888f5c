#!/bin/bash
# shared-memory carve-out of the physics kernel (percent of the SM's 228 KB; the rest is L1)
mkdir -p gpurun_out; rm -f gpurun_out/carve.txt
for c in default 30 40 60; do
  echo "== physics carveout $c" | tee -a gpurun_out/carve.txt
  if [ "$c" = default ]; then unset TDE_PHYS_CARVEOUT; else export TDE_PHYS_CARVEOUT=$c; fi
  python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/carve.txt
done
