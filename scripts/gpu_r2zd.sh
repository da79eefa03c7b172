#!/bin/bash
K3_RUNS=20 timeout 120 python scripts/k3_determinism.py 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_tc.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -4
for v in 4736 200000 2000000; do K3_V=$v K3_ITERS=6 timeout 120 python scripts/prof_k3.py 2>&1 | tail -1; done
AAE_B200_LIB=$PWD/aae-recommender_b200/build/variants/lib_k3x_trace.so timeout 120 python scripts/k3_trace.py 2>&1 | tail -2
K3_ITERS=1500 timeout 120 python scripts/k3_sustained.py 2>&1 | tail -2 | cut -c1-300
