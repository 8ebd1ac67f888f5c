#!/bin/bash
# Round 2, call N: the host expansion on a sleeping thread pool (one job per step) instead of OpenMP regions per chunk.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -k "compact_host or host_buffer" 2>&1 | tail -3
timeout 900 python tools/e2e_ab.py --threads 4,8,12,16,32 --chunks 0,8 --out gpurun_out/e2e_ab_n1_pool.json 2>&1 | tail -20
timeout 600 python tools/e2e_ab.py --envs 4096 --threads 8,16 --out gpurun_out/e2e_ab_n1_4096_pool.json 2>&1 | tail -8
timeout 600 python tools/e2e_ab.py --envs 1024 --threads 4,16 --out gpurun_out/e2e_ab_n1_1024_pool.json 2>&1 | tail -8
