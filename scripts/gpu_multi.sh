#!/bin/bash
# 2-GPU round trip: the whole GPU test suite (incl. the 2-GPU test), the sharded parity check, bench at N=2 (both arms)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n2_pytest.log; tail -4 gpurun_out/n2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "multi_check exit $?" >> gpurun_out/multi_check.log
tail -12 gpurun_out/multi_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 exit $?"
tail -c 600 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1])
print("N=2 MPD value %.0f e2e %.0f ms %.4f sustained %.4f hot %s" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['ms_per_step'], d.get('w1_all_rows_hot')))
print("parity", {k:v for k,v in d['parity_check'].items() if k!='tolerance'})
for k in ("mpd_b1000","mpd_b10000","pubmed","pubmed_b500","pubmed_cond"):
    x=d.get(k)
    if x: print(k, "value %.0f ms %.3f" % (x['value'], x['ms_per_step']))
print("sweep", {k:(round(v['value']),round(v['e2e'])) for k,v in d.get("mpd_predict_sweep",{}).items()})
print("set-sharded", {k:(round(v['value']),round(v['e2e'])) for k,v in d.get("mpd_predict_sweep_set_sharded",{}).items()})
print(d['roofline'].get('step_timeline_us'))
PY
