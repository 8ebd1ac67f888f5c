#!/bin/bash
# Round 2, call M (8 GPUs): the host-buffer step at N=8, RGB planes over PCIe against the class image + host expansion.
set -x
mkdir -p gpurun_out
nproc; nvidia-smi topo -m | head -12
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -k "compact_host" 2>&1 | tail -3
for th in default 2 8; do
  if [ "$th" = default ]; then unset TDE_HOST_THREADS; else export TDE_HOST_THREADS=$th; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench_n8_th$th.json 2> gpurun_out/r2m_bench_n8_th$th.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2m_bench_n8_th$th.json").read().strip().splitlines()[-1])
print("threads $th:", d["value"], d["e2e"])
PY
done
