#!/bin/bash
# A/B of CUDA library variants with a parity gate: tools/ab_checked.sh name1 name2 ...   ("base" = the in-tree library)
# Each variant first runs the parity tests (bit-exact against the oracle and the golden fixtures), then is timed.
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/ab.txt
  if [ "$v" = base ]; then unset TDE_B200_LIB; else export TDE_B200_LIB=$PWD/variants/lib_$v.so; fi
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_golden.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/ab.txt
  python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
done
