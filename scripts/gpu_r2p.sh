#!/bin/bash
# round 2 profiling run (1 GPU): launch list of the train step + ncu --set full captures of the kernels VERDICT r1 asked for
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# (1) launch list of the PubMed-shaped step (graph replay: kernel nodes are profiled individually)
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2p_launches_pubmed.csv \
  python bench.py --workload pubmed --no-extra --no-cpu --steps 3 --warmup 3 > gpurun_out/r2p_launches_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r2p_launches_pubmed.csv > gpurun_out/r2p_launches_pubmed_summary.txt; head -30 gpurun_out/r2p_launches_pubmed_summary.txt
# (2) K3 full capture
K3_ITERS=3 $NCU --set full --import-source on -k regex:dec_out_train_tc2 --launch-skip 1 -c 1 -f -o gpurun_out/r2p_k3 python scripts/prof_k3.py > gpurun_out/r2p_k3.log 2>&1
# (3) K5 v2 filter + candidate kernels (MPD shape, B=1000)
P_ITERS=1 $NCU --set full --import-source on --profile-from-start off -k regex:"dec_out_select2|cand_finish|cand_sort_small|row_kth" -c 5 -f -o gpurun_out/r2p_k5 python scripts/prof_predict.py > gpurun_out/r2p_k5.log 2>&1
# (4) the W1 kernels + batch kernels of the train step (eager launches)
$NCU --set full --import-source on -k regex:"w1_rows_update|w1_sweep_blocked|w1_catchup|batch_prepare|batch_gather" --launch-skip 20 -c 6 -f -o gpurun_out/r2p_w1 \
  python bench.py --workload pubmed --no-extra --no-cpu --no-graph --steps 3 --warmup 3 > gpurun_out/r2p_w1.log 2>&1
ls -la gpurun_out/r2p_*.ncu-rep
