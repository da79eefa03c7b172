#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k large_vocab 2>&1 | grep -E "^E  |passed|failed" | head -6; done
K3_V=47360 AAE_B200_LIB=$PWD/aae-recommender_b200/build/variants/lib_k3x_trace.so python scripts/k3_trace.py 2>&1 | tail -1
