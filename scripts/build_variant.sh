#!/bin/bash
# build_variant.sh NAME "<extra nvcc -D flags>": a copy of the library with experiment macros, under build/variants/
set -e
cd "$(dirname "$0")/../aae-recommender_b200/csrc"
NAME=$1; shift
OUT=../build/variants; mkdir -p $OUT/$NAME
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr"
for f in api bag w1_blocked mlp dec_out_simt dec_out_tc dec_out_select2 topk peer; do nvcc $FLAGS "$@" -c $f.cu -o $OUT/$NAME/$f.o & done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/lib_$NAME.so $OUT/$NAME/*.o
echo built $OUT/lib_$NAME.so
