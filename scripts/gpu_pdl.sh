#!/bin/bash
# programmatic dependent launch along the step's kernel chain (AAE_B200_PDL=1): parity tests, then the bench legs both ways
mkdir -p gpurun_out
AAE_B200_PDL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_tc.py tests/test_gpu_siblings.py -q -m gpu -x -n 3 -p no:cacheprovider > gpurun_out/pdl_pytest.log 2>&1; echo "pdl pytest rc=$?"; tail -5 gpurun_out/pdl_pytest.log
for v in 0 1 0 1; do
  AAE_B200_PDL=$v timeout 600 python bench.py --no-cpu --skip-big --steps 20 --warmup 5 > gpurun_out/pdl_bench_$v.json 2> gpurun_out/pdl_bench_$v.err; echo "bench rc=$?"; tail -2 gpurun_out/pdl_bench_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/pdl_bench_$v.json'))
print("== PDL=$v MPD value %.0f e2e %.0f ms %.4f sustained %.4f" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['ms_per_step']))
for k in ("mpd_b1000","pubmed","pubmed_b500","pubmed_cond","econbiz"):
    x=d.get(k)
    if x: print(k, "value %.0f ms %.4f" % (x['value'], x['ms_per_step']), ("e2e %.0f" % x['e2e']['value']) if 'e2e' in x else "", ("sustained %.4f" % x['sustained']['ms_per_step']) if x.get('sustained') else "")
print("timeline", d['pubmed']['roofline'].get('step_timeline_us'))
print("fit_epoch", d.get('fit_epoch',{}).get('ms_per_step_wall'))
PY
done 2>&1 | tee gpurun_out/pdl_summary.txt
