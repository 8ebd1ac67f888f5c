#!/bin/bash
# Round 2, call AK: C4 offroad against the number of grid cells.
set -x
for v in base cells65536 cells131072; do
  if [ "$v" = base ]; then unset TDE_B200_LIB; else export TDE_B200_LIB=$PWD/variants/lib_$v.so; fi
  echo "== $v"; python tools/c4_times.py | cut -c1-330
done
