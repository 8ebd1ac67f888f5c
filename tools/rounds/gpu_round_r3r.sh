#!/bin/bash
# Round 2, call AR: a 4 KB bitmap of SAFE cells in front of the 262 KB cell-record table (physics, C4 offroad).
set -x
tools/ab_checked.sh base safebits base safebits
for v in base safebits; do
  if [ "$v" = base ]; then unset TDE_B200_LIB; else export TDE_B200_LIB=$PWD/variants/lib_$v.so; fi
  echo "== $v"; python tools/c4_times.py | cut -c1-90; python tools/kernel_times.py 8192 8 | head -1
done
