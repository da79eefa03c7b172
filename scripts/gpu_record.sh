#!/bin/bash
# final N=1 record: tests, smoke, both bench arms as the driver runs them
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/rec_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/rec_pytest.log; tail -4 gpurun_out/rec_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/rec_smoke.log
timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/rec_bench_ref.json 2> gpurun_out/rec_bench_ref.err; echo "ref rc=$?"
timeout 2400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/rec_bench.json 2> gpurun_out/rec_bench.err; echo "bench rc=$?"; tail -c 500 gpurun_out/rec_bench.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/rec_bench_ref.json'))
print("REF value %.1f sets/s ms %.1f steps_timed %s cores %s kind %s" % (r['value'], r['ms_per_step'], r.get('steps_timed'), r['cpu_baseline']['cores'], r['cpu_baseline']['kind']))
d=json.load(open('gpurun_out/rec_bench.json'))
print("MPD value %.0f e2e %.0f ms %.4f sustained %.4f K3 ms %.3f frac %.3f step_frac %.3f clocks %s" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['ms'], d['roofline']['frac'], d['roofline']['step_frac'], d['clocks']))
for k in ("mpd_b1000","mpd_b10000","pubmed","pubmed_b500","pubmed_cond","econbiz"):
    x=d.get(k)
    if x: print(k, "value %.0f ms %.3f tensor_frac %.4f" % (x['value'], x['ms_per_step'], x['tensor_frac']), ("e2e %.0f"%x['e2e']['value']) if 'e2e' in x else "", ("K3 frac %.3f"%x['roofline']['frac']) if 'roofline' in x else "")
print("sweep", {k:(round(v['value']),round(v['e2e']), round(v['tensor_frac'],3)) for k,v in d.get("mpd_predict_sweep",{}).items()})
print("pubmed predict", round(d["pubmed"]["predict"]["value"]), round(d["pubmed"]["predict"]["e2e"]["value"]))
print("fit_epoch", d.get("fit_epoch",{}).get("ms_per_step_wall"), "cpu", d.get("cpu_baseline",{}).get("value"), "gpu_baseline", d.get("gpu_baseline",{}).get("value"))
PY
