#!/bin/bash
# build_variant.sh NAME [extra nvcc flags]: compile the library into build/libklp_NAME.so (A/B experiments; see KLP_LIBRARY)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off -shared "$@" -o build/libklp_$name.so lineprofiles_b200/csrc/klp_lib.cu
