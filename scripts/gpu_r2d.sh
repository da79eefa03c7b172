#!/bin/bash
mkdir -p gpurun_out
for B in 500 1000; do for seed in 21 22 23 24 25 26; do
  BATCH=$B SEED=$seed REPS=8 timeout 600 python scripts/dbg_race.py > gpurun_out/r2d_seed.log 2>&1; echo "B=$B seed=$seed BAD: $(grep -c BAD gpurun_out/r2d_seed.log) worst: $(grep -h 'rep' gpurun_out/r2d_seed.log | awk '{print $NF}' | sort -g | tail -1)"
done; done
