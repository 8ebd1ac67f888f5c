#!/bin/bash
# A/B timing of CUDA library variants built by tools/build_variant.sh: tools/ab_variants.sh name1 name2 ...
# (per-kernel device times of the C3 step; the in-tree library runs the GPU tests first)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/ab.txt
  TDE_B200_LIB=$PWD/variants/lib_$v.so python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
done
