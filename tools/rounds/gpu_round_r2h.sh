#!/bin/bash
# Round 2, call H: per-cell summary table on the global physics path (A/B), C5 with the slimmer collector loop, C4.
set -x
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
for v in "TDE_CELL16=1" "TDE_CELL16=0"; do
  echo "== $v" | tee -a gpurun_out/ab.txt
  env $v python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
  env $v python tools/c4_times.py 2>&1 | tail -3 | tee -a gpurun_out/ab.txt
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_h.txt
python bench.py --workload c5 --steps 256 --warmup 32 --no-cpu-baseline | tee gpurun_out/bench_c5.json
