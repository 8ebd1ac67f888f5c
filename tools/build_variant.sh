#!/bin/bash
# build an A/B variant of the CUDA library: tools/build_variant.sh <name> [-DFOO=1 ...]  -> gpurun_variants/lib_<name>.so
set -e
name=$1; shift
mkdir -p variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC -shared "$@" -o variants/lib_$name.so torchdriveenv_b200/csrc/tde_b200.cu
echo built variants/lib_$name.so
