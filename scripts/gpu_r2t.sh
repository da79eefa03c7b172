#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -25 gpurun_out/r2t_pytest.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
