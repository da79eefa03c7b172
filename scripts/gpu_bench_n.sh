#!/bin/bash
# bench at N GPUs (N = first argument), as the driver launches it
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus $N --steps 20 --warmup 5 $BENCH_ARGS > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench$N exit $?"
tail -c 400 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
print("N=$N MPD value %.0f e2e %.0f ms %.4f sustained %.4f hot %s" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['ms_per_step'], d.get('w1_all_rows_hot')))
print("parity", {k:v for k,v in d['parity_check'].items() if k!='tolerance'})
for k in ("mpd_b1000","mpd_b10000","pubmed","pubmed_b500","pubmed_cond"):
    x=d.get(k)
    if x: print(k, "value %.0f ms %.3f" % (x['value'], x['ms_per_step']))
print("sweep", {k:(round(v['value']),round(v['e2e'])) for k,v in d.get("mpd_predict_sweep",{}).items()})
print("set-sharded", {k:(round(v['value']),round(v['e2e'])) for k,v in d.get("mpd_predict_sweep_set_sharded",{}).items()})
print(d['roofline'].get('step_timeline_us'))
PY
