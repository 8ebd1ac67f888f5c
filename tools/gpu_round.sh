#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), ncu launch list + full capture of the step kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
python bench.py 2> gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --policy random --e2e-steps 3 > gpurun_out/ncu_bench1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tde_.*_kernel -s 13 -c 2 -o gpurun_out/prof_step -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-other-configs --policy random --e2e-steps 3 > gpurun_out/ncu_bench2.log 2>&1
ls -la gpurun_out | tail -12
