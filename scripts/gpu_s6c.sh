#!/bin/bash
# evidence run at HEAD: bench (both arms), ncu launch list of the train step, full captures of K3 and of the K5 select kernel
mkdir -p gpurun_out
timeout 900 python bench.py --steps 200 --warmup 10 --kernel-times > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log; tail -2 gpurun_out/bench.log | cut -c1-600
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-graph --no-extra > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_out_train_tc2 -s 2 -c 1 -f -o gpurun_out/prof_k3_s6 \
  python scripts/prof_k3.py > gpurun_out/ncu_full_k3.log 2>&1; echo "ncu k3 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_out_select -s 3 -c 1 -f -o gpurun_out/prof_k5_s6 \
  python scripts/prof_predict.py > gpurun_out/ncu_full_k5.log 2>&1; echo "ncu k5 exit $?"
ls -la gpurun_out/*.ncu-rep
