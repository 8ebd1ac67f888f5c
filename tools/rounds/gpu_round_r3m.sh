#!/bin/bash
# Round 2, call AM: physics tickets drawn mid-env (pt4), render ticket published at the end of the env (rlate), both.
set -x
tools/ab_checked.sh base pt4 rlate both base
for v in base both; do
  if [ "$v" = base ]; then unset TDE_B200_LIB; else export TDE_B200_LIB=$PWD/variants/lib_$v.so; fi
  echo "== $v"; python tools/kernel_times.py 8192 8 | head -1; python tools/kernel_times.py 1024 16 | head -1
done
