#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q --timeout=120 -x -k "operand_views" > gpurun_out/tc_selftest.log 2>&1; echo "selftest exit $?" >> gpurun_out/tc_selftest.log
tail -25 gpurun_out/tc_selftest.log
timeout 600 python -m pytest tests/test_gpu_tc.py -q --timeout=300 -k "not operand_views" > gpurun_out/tc_tests.log 2>&1; echo "tc tests exit $?" >> gpurun_out/tc_tests.log
tail -30 gpurun_out/tc_tests.log
