#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_multi.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -3
GS=32 timeout 300 python scripts/g_probe.py 2>&1 | tail -1
AAE_B200_SWEEP_CTAS=8 GS=32,16 timeout 300 python scripts/g_probe.py 2>&1 | tail -2
