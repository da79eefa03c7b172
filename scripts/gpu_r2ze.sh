#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r2ze_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2ze_pytest.log; tail -4 gpurun_out/r2ze_pytest.log
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --skip-big > gpurun_out/r2ze_bench.json 2> gpurun_out/r2ze_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2ze_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ze_bench.json'))
print("MPD value %.0f e2e %.0f ms %.4f sustained %.4f K3 ms %.3f frac %.3f step_frac %.3f clocks %s" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['ms'], d['roofline']['frac'], d['roofline']['step_frac'], d['clocks']))
x=d.get('pubmed')
if x: print("pubmed value %.0f ms %.3f K3 frac %.3f ms %.4f" % (x['value'], x['ms_per_step'], x['roofline']['frac'], x['roofline']['ms']))
print(d.get('parity_check'))
PY
