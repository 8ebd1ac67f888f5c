#!/bin/bash
# Round 2, call AX: AVX-512 against AVX2 in the host expansion, on a fresh box and after the GPU test suite has run on it.
set -x
grep -o "avx512bw" /proc/cpuinfo | head -1
python tools/e2e_ab.py --threads 8,16 --simd avx2,avx512 --steps 20 2>&1 | cut -c1-190
python -m pytest tests -x -q -m gpu 2>&1 | tail -1
python tools/e2e_ab.py --threads 16 --simd avx2,avx512 --steps 20 2>&1 | cut -c1-190
