#!/bin/bash
# final profiling run (1 GPU): launch list of the MPD-shaped train step, ncu --set full of the decoder kernel (both shapes) and of the W1 sweep
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/prof_launches_mpd.csv \
  python bench.py --no-extra --no-cpu --steps 3 --warmup 3 > gpurun_out/prof_launches_bench.log 2>&1; echo "launch list exit $?"
python scripts/launch_summary.py gpurun_out/prof_launches_mpd.csv > gpurun_out/prof_launches_mpd_summary.txt; head -24 gpurun_out/prof_launches_mpd_summary.txt
K3_ITERS=3 timeout 600 $NCU --set full --import-source on -k regex:dec_out_train_tc2 --launch-skip 1 -c 1 -f -o gpurun_out/prof_k3_pubmed python scripts/prof_k3.py > gpurun_out/prof_k3a.log 2>&1; echo "k3 pubmed exit $?"
K3_V=2000000 K3_ITERS=3 timeout 600 $NCU --set full --import-source on -k regex:dec_out_train_tc2 --launch-skip 1 -c 1 -f -o gpurun_out/prof_k3_mpd python scripts/prof_k3.py > gpurun_out/prof_k3b.log 2>&1; echo "k3 mpd exit $?"
timeout 600 $NCU --set full --import-source on -k regex:"w1_sweep_blocked" --launch-skip 40 -c 1 -f -o gpurun_out/prof_sweep \
  python bench.py --no-extra --no-cpu --no-graph --steps 3 --warmup 3 > gpurun_out/prof_sweep.log 2>&1; echo "sweep exit $?"
ls -la gpurun_out/prof_*.ncu-rep
