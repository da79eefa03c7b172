#!/bin/bash
# r3a: cold-row W1 sweep (product lib) + K3 with TMA bulk stores in E2 (variant lib): parity, kernel time, step time
mkdir -p gpurun_out
VL=$PWD/aae-recommender_b200/build/variants/lib_k3bulk.so
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "cold_rows or large_vocab or twenty_steps or graph_replay" -p no:cacheprovider > gpurun_out/r3a_pytest_base.log 2>&1; echo "base pytest rc=$?"; tail -3 gpurun_out/r3a_pytest_base.log
AAE_B200_LIB=$VL timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_configs.py tests/test_gpu_parity.py -q -m gpu -x -n 3 -p no:cacheprovider > gpurun_out/r3a_pytest_bulk.log 2>&1; echo "bulk pytest rc=$?"; tail -3 gpurun_out/r3a_pytest_bulk.log
for lib in base bulk; do
  if [ $lib = bulk ]; then export AAE_B200_LIB=$VL; else unset AAE_B200_LIB; fi
  echo "== $lib"
  K3_ITERS=12 timeout 120 python scripts/prof_k3.py 2>&1 | tail -1
  K3_V=2000000 K3_ITERS=12 timeout 120 python scripts/prof_k3.py 2>&1 | tail -1
  K3_ITERS=1500 timeout 120 python scripts/k3_sustained.py 2>&1 | tail -2
  timeout 300 python bench.py --no-extra --no-cpu --steps 50 --warmup 5 > gpurun_out/r3a_bench_$lib.json 2> gpurun_out/r3a_bench_$lib.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r3a_bench_$lib.json'))
print("MPD value %.0f e2e %.0f ms %.4f sustained %.4f K3 ms %.3f frac %.3f" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['ms'], d['roofline']['frac']))
print(d['roofline'].get('step_timeline_us'))
PY
done 2>&1 | tee gpurun_out/r3a_summary.txt
