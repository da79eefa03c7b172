#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_predict_fused.py tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2f_fused.log 2>&1
echo "rc=$?"; tail -12 gpurun_out/r2f_fused.log | cut -c1-300
timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu --skip-big > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print("sweep", json.dumps(d.get("mpd_predict_sweep"), indent=0))
print("pubmed predict", d["pubmed"]["predict"]["value"], d["pubmed"]["predict"]["path"], d["pubmed"]["predict"]["e2e"]["value"])
PY
