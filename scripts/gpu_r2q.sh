#!/bin/bash
for B in 1000 4000; do for cy in 1 2 4 8; do
echo "== B=$B CY=$cy"; AAE_B200_K5_CLUSTER=$cy P_V=2000000 P_B=$B P_ITERS=5 python scripts/prof_predict.py 2>&1 | grep -E "topk2|tail" 
done; done
