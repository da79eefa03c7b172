#!/bin/bash
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2zl_bench.json 2> gpurun_out/r2zl.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2zl_bench.json'))
print("MPD value %.0f e2e %.0f (%.3f ms) ms %.4f sustained %.4f K3 ms %.3f frac %.3f hot %s" % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['ms'], d['roofline']['frac'], d['roofline'].get('after_sustained')))
print(d['roofline'].get('step_timeline_us')); print(d['clocks'], d['roofline']['step_frac'], d['roofline']['step_frac_sustained'])
PY
