#!/bin/bash
K3_RUNS=30 timeout 120 python scripts/k3_determinism.py 2>&1 | tail -4
K3_V=2000000 K3_RUNS=8 timeout 120 python scripts/k3_determinism.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_tc.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -4
for v in 4736 200000 2000000; do K3_V=$v K3_ITERS=6 timeout 120 python scripts/prof_k3.py 2>&1 | tail -1; done
AAE_B200_LIB=$PWD/aae-recommender_b200/build/variants/lib_k3x_trace.so timeout 120 python scripts/k3_trace.py 2>&1 | tail -2
SUSTAIN_S=2 timeout 200 python scripts/step_trace.py 2>&1 | tail -4 | cut -c1-330
