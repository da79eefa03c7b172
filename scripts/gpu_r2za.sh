#!/bin/bash
mkdir -p gpurun_out
VD=$PWD/aae-recommender_b200/build/variants
for n in ${K3T_VARIANTS:-trace trace_nosimt trace_nomma}; do
  echo "=== $n"
  AAE_B200_LIB=$VD/lib_k3x_$n.so python scripts/k3_trace.py 2>&1 | tail -36
done | tee gpurun_out/k3_trace.txt | grep -v "^[0-9]" 
