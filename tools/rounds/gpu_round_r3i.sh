#!/bin/bash
# Round 2, call AI: tile size of the static render index (primitives up to 16 m stay in the tiles).
set -x
for t in 8 6 4 3 2 8 4; do echo "tile $t"; TDE_TILE_M=$t python tools/kernel_times.py 2>&1 | head -1; done
TDE_TILE_M=4 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_golden.py -x -q 2>&1 | tail -1
for t in 8 4; do echo "tile $t"; TDE_TILE_M=$t python tools/kernel_times.py 8192 8 2>&1 | head -1; done
