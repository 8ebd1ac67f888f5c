#!/bin/bash
# Round 2, call AJ: cells of the lane-mesh grid of the physics kernel (8 B per cell; 32,768 cells = 262 KB at C3).
set -x
tools/ab_checked.sh base cells8192 cells16384 cells65536 base
for v in base cells8192 cells16384; do
  if [ "$v" = base ]; then unset TDE_B200_LIB; else export TDE_B200_LIB=$PWD/variants/lib_$v.so; fi
  echo "== $v"; python tools/c4_times.py | cut -c1-90; python tools/kernel_times.py 8192 8 | head -1
done
