#!/bin/bash
# Build an alternative libxde (kernel tuning experiments; timed by tools/adj_variants.py on the GPU box).
# usage: tools/build_variant.sh NAME file.cu "-DFOO=1 -DBAR=2"   ->  tools/_variants/libxde_NAME.so
set -e
cd "$(dirname "$0")/../paddlexde_b200/csrc"
NAME=$1; SRC=$2; DEFS=$3
OUT=../../tools/_variants
mkdir -p $OUT/obj_$NAME
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -std=c++17 -lineinfo -fmad=false $ARCH -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I../../include -I. \
  --expt-relaxed-constexpr $DEFS -c $SRC -o $OUT/obj_$NAME/${SRC%.cu}.o
OBJS=$(ls build/*.o | grep -v "build/${SRC%.cu}.o")
$NVCC $ARCH -shared -cudart static -o $OUT/libxde_$NAME.so $OBJS $OUT/obj_$NAME/${SRC%.cu}.o
echo built $OUT/libxde_$NAME.so
