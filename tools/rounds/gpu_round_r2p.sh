#!/bin/bash
# Round 2, call P: the fused step (one launch: physics + render by the same warp) against the two-launch step.
set -x
mkdir -p gpurun_out
TDE_FUSED_STEP=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_golden.py tests/test_gpu_round2.py -x -q 2>&1 | tail -4
for f in 0 1 0 1; do
  TDE_FUSED_STEP=$f python tools/kernel_times.py | head -1
done
TDE_FUSED_STEP=1 python tools/kernel_times.py 8192 8 | head -1
TDE_FUSED_STEP=0 python tools/kernel_times.py 8192 8 | head -1
TDE_FUSED_STEP=1 python tools/kernel_times.py 1024 16 | head -1
TDE_FUSED_STEP=0 python tools/kernel_times.py 1024 16 | head -1
