#!/bin/bash
VD=$PWD/aae-recommender_b200/build/variants
for n in base nomma nosimt noe2 noe2st noe1math nog1 nog2 nog3 nowt; do
  echo -n "== $n: "; AAE_B200_LIB=$VD/lib_k3x_$n.so K3_ITERS=2000 python scripts/k3_sustained.py 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-330; echo
done | tee gpurun_out/k3_sustained_variants.txt
