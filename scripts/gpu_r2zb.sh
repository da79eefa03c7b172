#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_tc.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -5
K3_ITERS=8 python scripts/prof_k3.py 2>&1 | tail -1
K3_V=2000000 K3_ITERS=6 python scripts/prof_k3.py 2>&1 | tail -1
AAE_B200_LIB=$PWD/aae-recommender_b200/build/variants/lib_k3x_trace.so python scripts/k3_trace.py 2>&1 | tail -4 | tee gpurun_out/k3_trace2.txt
