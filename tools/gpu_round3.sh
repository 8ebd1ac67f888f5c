#!/bin/bash
# GPU tests, then parity-gated A/B timing of the named library variants ("base" = in-tree)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
bash tools/ab_checked.sh "$@"
