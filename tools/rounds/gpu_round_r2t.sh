#!/bin/bash
# Round 2, call T: render envs handed out most expensive first (lists filed by the previous launch) against index order.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_golden.py tests/test_gpu_round2.py tests/test_gpu_env_api.py -x -q 2>&1 | tail -3
for o in 1 0 1 0; do
  TDE_RENDER_ORDER=$o python tools/kernel_times.py 2>&1 | head -1
done
TDE_RENDER_ORDER=1 python tools/kernel_times.py 8192 8 2>&1 | head -1
TDE_RENDER_ORDER=0 python tools/kernel_times.py 8192 8 2>&1 | head -1
tools/build_variant.sh trace -DTDE_TRACE > /dev/null 2>&1
TDE_B200_LIB=variants/lib_trace.so python tools/trace_envs.py 2>&1 | grep -A13 "^render"
