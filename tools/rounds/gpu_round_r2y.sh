#!/bin/bash
# Round 2, call Y: offroad kernel with the containment tests in the flattened pass and the owner search on shuffles; CTAs per SM.
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "c4 or offroad or collision" 2>&1 | tail -3
for c in 8 6 5 4 3; do TDE_OFFROAD_CTAS=$c python tools/c4_times.py | cut -c1-80; done
