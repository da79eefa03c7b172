#!/bin/bash
# One GPU-box round trip: parity tests, smoke, bench (both arms), ncu launch list and one full
# capture of the decoder-output kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 200 --warmup 10 --kernel-times ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log; tail -5 gpurun_out/bench.log; tail -30 gpurun_out/bench.err
if [ -n "${REF_ARM}" ]; then
  timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2>&1; tail -3 gpurun_out/bench_ref.log
fi
if [ -n "${NCU_LIST}" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-graph ${BENCH_ARGS} > gpurun_out/ncu_bench.log 2>&1
  echo "ncu list exit $?"
fi
if [ -n "${NCU_FULL}" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_FULL} -s 2 -c 2 -f -o gpurun_out/prof_${NCU_TAG:-k3} \
    python scripts/prof_k3.py > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
fi
