#!/bin/bash
mkdir -p gpurun_out
python scripts/dbg_bigbatch.py > gpurun_out/r2b_dbg.log 2>&1
tail -60 gpurun_out/r2b_dbg.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
