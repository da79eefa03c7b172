#!/bin/bash
# K3 iteration loop: tensor-core kernel tests, then the kernel's own timing on the PubMed shape.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q --timeout=300 -x > gpurun_out/tc_tests.log 2>&1; echo "tc tests exit $?" >> gpurun_out/tc_tests.log
tail -15 gpurun_out/tc_tests.log
for impl in ${K3_IMPLS:-1 3}; do K3_IMPL=$impl K3_ITERS=8 timeout 120 python scripts/prof_k3.py; done 2>&1 | tee gpurun_out/k3_times.log
