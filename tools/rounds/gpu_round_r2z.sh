#!/bin/bash
# Round 2, call Z: offroad kernel, CTAs per SM x shared-memory carve-out.
set -x
for c in 6 5 7; do for k in -1 auto 40 100; do
  if [ $k = auto ]; then unset TDE_OFFROAD_CARVE; else export TDE_OFFROAD_CARVE=$k; fi
  echo "ctas $c carve $k"; TDE_OFFROAD_CTAS=$c python tools/c4_times.py | cut -c1-80
done; done
