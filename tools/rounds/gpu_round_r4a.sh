#!/bin/bash
# Round 2, call BA: offroad kernel with its per-corner scratch laid out bank = lane (13 M bank conflicts per call before).
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -k "c4 or offroad or collision" 2>&1 | tail -1
for i in 1 2; do
python tools/c4_times.py | cut -c1-80
TDE_B200_LIB=$PWD/variants/lib_prev.so python tools/c4_times.py | cut -c1-80
done
