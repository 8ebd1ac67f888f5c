#!/bin/bash
# A/B timing only (no tests): tools/ab_quick.sh name1 name2 ...
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/ab.txt
  TDE_B200_LIB=$PWD/variants/lib_$v.so python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
done
