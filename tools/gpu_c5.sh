#!/bin/bash
# C5 round: rollout-collector test + the C5 / C2 / C4 bench lines
mkdir -p gpurun_out
python -m pytest tests/test_gpu_env_api.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_c5.txt
python bench.py --workload c5 --steps 256 --warmup 32 --cpu-seconds 6 2> gpurun_out/bench_c5_err.txt | tee gpurun_out/bench_c5.json
tail -5 gpurun_out/bench_c5_err.txt
python bench.py --workload c2 --steps 2000 --warmup 50 --cpu-seconds 4 2>/dev/null | tee gpurun_out/bench_c2.json
python bench.py --workload c4 --steps 50 --warmup 5 --cpu-seconds 4 2>/dev/null | tee gpurun_out/bench_c4.json
python tools/kernel_times.py 2>&1 | tail -2
