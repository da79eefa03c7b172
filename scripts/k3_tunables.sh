#!/bin/bash
# Compile-time tunables of the decoder-output training kernel, one variant library each (scripts/build_variant.sh NAME
# dec_out_tc.cu -D...): stand-alone kernel time at both shapes.  usage (GPU box): scripts/k3_tunables.sh name1 name2 ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VD=$PWD/aae-recommender_b200/build/variants
for name in base "$@" base; do
  if [ $name = base ]; then unset AAE_B200_LIB; else export AAE_B200_LIB=$VD/lib_$name.so; fi
  echo -n "$name | "; K3_ITERS=16 python scripts/prof_k3.py 2>&1 | tail -1 | sed 's/.*launch: //' | tr -d '\n'
  echo -n " | "; K3_V=2000000 K3_ITERS=16 python scripts/prof_k3.py 2>&1 | tail -1 | sed 's/.*launch: //'
done | tee gpurun_out/k3_tunables.txt
