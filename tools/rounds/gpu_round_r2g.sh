#!/bin/bash
# Round 2, call G: the GPU tests that changed, C5 with scatter vs shift, full default bench line.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_env_api.py tests/test_gpu_round2.py -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_g.txt
python bench.py --workload c5 --steps 256 --warmup 32 --no-cpu-baseline | tee gpurun_out/bench_c5.json
python bench.py --steps 2000 --warmup 50 2> gpurun_out/bench_err.txt | tee gpurun_out/bench.json
ls -la gpurun_out | tail -5
