#!/bin/bash
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-extra --kernel-times > gpurun_out/r2zi_bench.json 2> gpurun_out/r2zi_bench.err; echo rc=$?
grep -v "^$" gpurun_out/r2zi_bench.err | tail -30
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2zi_bench.json'))
print("MPD value %.0f e2e %.0f ms %.4f sustained %.4f K3 ms %.3f frac %.3f" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['ms'], d['roofline']['frac']))
PY
K3_V=2000000 K3_ITERS=6 timeout 120 python scripts/prof_k3.py 2>&1 | tail -1
