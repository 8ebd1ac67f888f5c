#!/bin/bash
# Round 2, call O (8 GPUs): the host-buffer step at N=8 with the thread pool.
set -x
mkdir -p gpurun_out
for th in default 6; do
  if [ "$th" = default ]; then unset TDE_HOST_THREADS; else export TDE_HOST_THREADS=$th; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_bench_n8_th$th.json 2> gpurun_out/r2o_bench_n8_th$th.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2o_bench_n8_th$th.json").read().strip().splitlines()[-1])
print("threads $th:", d["value"], d["e2e"])
PY
done
