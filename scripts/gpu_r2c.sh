#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider -s -k "mn_major" > gpurun_out/r2c_mn.log 2>&1
grep -n "MN-major\|passed\|failed" gpurun_out/r2c_mn.log | tail -5
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -k "script_batch" 2>&1 | tail -2; done
timeout 900 python bench.py --workload pubmed --no-extra --no-cpu --steps 50 --warmup 5 > gpurun_out/r2c_bench_pubmed.json 2> gpurun_out/r2c_bench_pubmed.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2c_bench_pubmed.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench_pubmed.json'))
print("pubmed value %.0f e2e %.0f ms %.4f sustained %.4f  K3 ms %.4f frac %.3f step_frac %.3f" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['ms'], d['roofline']['frac'], d['roofline']['step_frac']))
PY
