#!/bin/bash
# build an A/B variant of the library: scripts/build_variant.sh NAME -DFOO=1 ...  -> scripts/_var/NAME.so
# (run it with MYRRIX_ALS_LIB=scripts/_var/NAME.so; development only)
cd "$(dirname "$0")/.." && mkdir -p scripts/_var
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 "$@" -Xcompiler -fPIC -shared \
  -o scripts/_var/$name.so myrrix-recommender_b200/csrc/als_abi.cu -ldl
