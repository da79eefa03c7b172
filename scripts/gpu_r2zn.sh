#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -5
K3_RUNS=10 timeout 120 python scripts/k3_determinism.py 2>&1 | tail -2
timeout 400 python scripts/age_probe2.py 2>&1 | tail -13 | cut -c1-60,130-250
