#!/bin/bash
# build_variant.sh NAME FILE.cu "<extra nvcc -D flags>": a copy of the library in which ONE source file is compiled with
# experiment macros (the other objects are the product build's), under build/variants/lib_NAME.so.
# Use: AAE_B200_LIB=$PWD/aae-recommender_b200/build/variants/lib_NAME.so python ...
set -e
cd "$(dirname "$0")/../aae-recommender_b200/csrc"
NAME=$1; FILE=$2; shift 2
make -s all
OUT=../build/variants; mkdir -p $OUT
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr"
nvcc $FLAGS "$@" -c $FILE -o $OUT/${NAME}_${FILE%.cu}.o
OBJS=$(ls ../build/*.o | grep -v "/${FILE%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/lib_$NAME.so $OUT/${NAME}_${FILE%.cu}.o $OBJS
echo built $OUT/lib_$NAME.so
