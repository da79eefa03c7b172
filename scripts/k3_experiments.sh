#!/bin/bash
# Timing experiments on the decoder-output training kernel: variants of dec_out_tc.cu with parts switched off
# (K3X_* macros; results are WRONG by construction, only the time is read) to see what bounds the tile loop.
# usage: scripts/k3_experiments.sh build   (here, no GPU)   |   scripts/k3_experiments.sh run   (on the GPU box)
set -e
cd "$(dirname "$0")/.."
CS=aae-recommender_b200/csrc; BD=aae-recommender_b200/build; VD=$BD/variants
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr"
VARIANTS="base: noe2math:-DK3X_NO_E2MATH noe2st:-DK3X_NO_E2ST noe2ld:-DK3X_NO_E2LD noe2:-DK3X_NO_E2MATH,-DK3X_NO_E2ST,-DK3X_NO_E2LD noe1math:-DK3X_NO_E1MATH nog1:-DK3X_NO_G1 nog2:-DK3X_NO_G2 nog3:-DK3X_NO_G3 nomma:-DK3X_NO_G1,-DK3X_NO_G2,-DK3X_NO_G3 nowt:-DK3X_NO_WT nosimt:-DK3X_NO_E2MATH,-DK3X_NO_E2ST,-DK3X_NO_E2LD,-DK3X_NO_E1MATH,-DK3X_NO_WT"
if [ "$1" = "build" ]; then
  mkdir -p $VD
  for v in $VARIANTS; do
    name=${v%%:*}; defs=$(echo ${v#*:} | tr ',' ' ')
    ( nvcc $FLAGS $defs -c $CS/dec_out_tc.cu -o $VD/k3x_$name.o && \
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $VD/lib_k3x_$name.so $VD/k3x_$name.o \
        $(ls $BD/*.o | grep -v /dec_out_tc.o) && echo built $name ) &
  done
  wait
else
  mkdir -p gpurun_out
  for v in $VARIANTS; do
    name=${v%%:*}
    echo -n "$name: "
    AAE_B200_LIB=$PWD/$VD/lib_k3x_$name.so K3_ITERS=8 python scripts/prof_k3.py 2>&1 | tail -1
  done | tee gpurun_out/k3_experiments.txt
fi
