#!/bin/bash
# Builds an A/B variant of librtb.so with extra nvcc flags: tools/build_variant.sh NAME "-DRTB_LEAF_MAX=2 ..."
# -> rust_raytrace_b200/csrc/build/variants/librtb_NAME.so, selected at run time with RTB_LIB=<path>.
set -eu
NAME=$1; FLAGS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
V=$ROOT/rust_raytrace_b200/csrc/build/variants
mkdir -p $V/obj_$NAME
cd $ROOT/rust_raytrace_b200/csrc
for f in rtb_api rtb_lbvh rtb_scene rtb_ext rtb_trace rtb_wavefront; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math \
       -I$ROOT/include -I. $FLAGS -c $f.cu -o $V/obj_$NAME/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/librtb_$NAME.so $V/obj_$NAME/*.o build/raytrace_host.o -lcudart_static -lpthread -ldl -lrt
echo $V/librtb_$NAME.so
