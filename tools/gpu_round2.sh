#!/bin/bash
# One gpurun call: GPU tests, bench (e2e through the chunked host path), A/B kernel variants, compute-sanitizer on a small run.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
python bench.py --steps 100 --warmup 10 2> gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
rm -f gpurun_out/ab.txt
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/ab.txt
  if [ "$v" = base ]; then python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
  else TDE_B200_LIB=$PWD/variants/lib_$v.so python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt; fi
done
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.txt 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_summary.txt
  tail -3 gpurun_out/sanitize_$tool.txt
done
