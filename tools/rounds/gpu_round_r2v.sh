#!/bin/bash
# Round 2, call V: the C4 offroad kernel with the (corner, candidate) pairs flattened, against the previous build.
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "c4 or offroad or collision" 2>&1 | tail -3
python tools/c4_times.py
TDE_B200_LIB=$PWD/variants/lib_prev.so python tools/c4_times.py
python tools/c4_times.py 100
TDE_B200_LIB=$PWD/variants/lib_prev.so python tools/c4_times.py 100
