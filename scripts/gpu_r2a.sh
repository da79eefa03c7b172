#!/bin/bash
# round 2, run A: full GPU test suite + default bench (1 GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -30 gpurun_out/r2a_pytest.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.err
head -c 3000 gpurun_out/r2a_bench.json
