#!/bin/bash
# Final evidence run (1 GPU): parity tests, smoke, bench (both arms), ncu launch list, full captures of K3 and K5.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -q -m gpu -n 4 --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 200 --warmup 10 --kernel-times > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log; tail -2 gpurun_out/bench.log | cut -c1-400
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-graph --no-extra > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_out_train_tc2 -s 2 -c 1 -f -o gpurun_out/prof_k3_s6 \
  python scripts/prof_k3.py > gpurun_out/ncu_full_k3.log 2>&1; echo "ncu k3 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_out_select -s 3 -c 1 -f -o gpurun_out/prof_k5_s6 \
  python scripts/prof_predict.py > gpurun_out/ncu_full_k5.log 2>&1; echo "ncu k5 exit $?"
P_ITERS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/predict_launches_mpd.csv python scripts/prof_predict.py > gpurun_out/ncu_predict.log 2>&1
P_V=200000 P_ITERS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/predict_launches_pubmed.csv python scripts/prof_predict.py >> gpurun_out/ncu_predict.log 2>&1
