#!/bin/bash
for idle in 0 1.0; do
AAE_BENCH_IDLE_S=$idle timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2zl_bench_$idle.json 2> gpurun_out/r2zl.err; echo "idle $idle rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2zl_bench_$idle.json'))
print("MPD value %.0f e2e %.0f (%.3f ms) ms %.4f sustained %.4f K3 ms %.3f frac %.3f hot %.3f" % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['ms'], d['roofline']['frac'], d['roofline']['after_sustained']['ms']))
PY
done
