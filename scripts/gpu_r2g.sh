#!/bin/bash
mkdir -p gpurun_out
for cfg in "2000000 1000" "200000 1000"; do set -- $cfg
P_V=$1 P_B=$2 P_ITERS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file gpurun_out/r2g_predict_launches_$1_$2.csv python scripts/prof_predict.py > gpurun_out/r2g_prof_$1_$2.log 2>&1
echo "== V=$1 B=$2"; python scripts/launch_summary.py gpurun_out/r2g_predict_launches_$1_$2.csv | head -16
done
P_V=2000000 P_B=1000 P_ITERS=5 python scripts/prof_predict.py 2>&1 | tail -3
P_V=2000000 P_B=4000 P_ITERS=5 python scripts/prof_predict.py 2>&1 | tail -3
