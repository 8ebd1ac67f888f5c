#!/bin/bash
# Round 2, call L: tde_step_host with the 4-bit class image over PCIe + host expansion against the RGB planes over PCIe.
set -x
mkdir -p gpurun_out
nproc; lscpu | grep -i "model name\|^CPU(s)\|socket\|numa" 
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "compact_host" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "step_host" 2>&1 | tail -3
timeout 900 python tools/e2e_ab.py --threads 2,4,8,16 --chunks 0,8 --out gpurun_out/e2e_ab_n1.json 2>&1 | tail -20
timeout 600 python tools/e2e_ab.py --envs 4096 --threads 8,16 --out gpurun_out/e2e_ab_n1_4096.json 2>&1 | tail -8
