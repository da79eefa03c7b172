#!/bin/bash
# Session-6 GPU round trip (1 GPU): K3 prefetch variants, parity tests, smoke, bench with predict + MPD legs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for v in "" pf0 pf3 pf6; do
  if [ -n "$v" ]; then export AAE_B200_LIB=$PWD/aae-recommender_b200/build/variants/lib_$v.so; else unset AAE_B200_LIB; fi
  echo "variant ${v:-default}" >> gpurun_out/k3_variants.log
  K3_ITERS=12 timeout 300 python scripts/prof_k3.py >> gpurun_out/k3_variants.log 2>&1
done
unset AAE_B200_LIB
cat gpurun_out/k3_variants.log
timeout 1200 python -m pytest tests -q -m gpu -n 4 --timeout=900 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 200 --warmup 10 --kernel-times ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log; tail -5 gpurun_out/bench.log; tail -25 gpurun_out/bench.err
