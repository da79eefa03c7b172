#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -n 4 --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
P_ITERS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/predict_launches_mpd.csv python scripts/prof_predict.py > gpurun_out/ncu_predict.log 2>&1
P_V=200000 P_ITERS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/predict_launches_pubmed.csv python scripts/prof_predict.py >> gpurun_out/ncu_predict.log 2>&1
python scripts/launch_summary.py gpurun_out/predict_launches_mpd.csv | head -12
python scripts/launch_summary.py gpurun_out/predict_launches_pubmed.csv | head -12
