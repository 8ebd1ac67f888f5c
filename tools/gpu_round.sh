#!/bin/bash
# One gpurun call: GPU tests, bench (both arms), staged-vs-default physics, ncu launch list + full capture of the step kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
python bench.py --steps 2000 --warmup 50 2> gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/bench_ref.json
rm -f gpurun_out/ab.txt
for v in "TDE_PHYS_STAGE=0" "TDE_PHYS_STAGE=1" "TDE_PHYS_STAGE=1 TDE_PHYS_WARPS=16"; do
  echo "== $v" | tee -a gpurun_out/ab.txt
  env $v python tools/kernel_times.py 2>&1 | head -1 | tee -a gpurun_out/ab.txt
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --policy random --e2e-steps 3 > gpurun_out/ncu_bench1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tde_.*_kernel -s 13 -c 2 -o gpurun_out/prof_step -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-other-configs --policy random --e2e-steps 3 > gpurun_out/ncu_bench2.log 2>&1
ls -la gpurun_out
