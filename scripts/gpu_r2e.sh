#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -8 gpurun_out/r2e_pytest.log
timeout 900 python bench.py --workload pubmed --no-extra --no-cpu --steps 50 --warmup 5 > gpurun_out/r2e_bench_pubmed.json 2> gpurun_out/r2e_bench_pubmed.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_pubmed.json'))
print("pubmed value %.0f e2e %.0f ms %.4f sustained %.4f  K3 ms %.4f frac %.3f step_frac %.3f" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['ms'], d['roofline']['frac'], d['roofline']['step_frac']))
PY
timeout 2400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
def g(k): 
    x=d.get(k)
    return x
print("MPD value %.0f e2e %.0f ms %.4f sustained %.4f K3 frac %.3f step_frac %.3f" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac']))
for k in ("mpd_b1000","mpd_b10000","pubmed","pubmed_b500","pubmed_cond","econbiz"):
    x=d.get(k)
    if x: print(k, "value %.0f ms %.3f kernel %s tensor_frac %.4f" % (x['value'], x['ms_per_step'], x['decoder_kernel'], x['tensor_frac']), "e2e %.0f"%x['e2e']['value'] if 'e2e' in x else "")
print("sweep", d.get("mpd_predict_sweep"))
print("pubmed predict", d["pubmed"]["predict"]["value"] if "pubmed" in d else None)
print("fit_epoch", d.get("fit_epoch"))
print("cpu", d.get("cpu_baseline")); print("gpu_baseline", d.get("gpu_baseline"))
PY
