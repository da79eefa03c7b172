#!/bin/bash
# N-GPU bench (N = $1)
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err; echo "bench exit $?"
tail -c 800 gpurun_out/r2n_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2n_bench_n$N.json') if l.startswith('{')][-1])
    print("N=$N MPD value %.0f e2e %.0f ms %.4f sustained %.4f" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['ms_per_step']))
    print("parity_check", d.get("parity_check"))
    for k in ("mpd_b1000","mpd_b10000","pubmed","pubmed_b500","pubmed_cond"):
        x=d.get(k)
        if x: print(k, "value %.0f ms %.3f" % (x['value'], x['ms_per_step']))
    print("sweep items", {k:(round(v['value']),v['path']) for k,v in d.get("mpd_predict_sweep",{}).items()})
    print("sweep sets ", {k:(round(v['value']),v['path']) for k,v in d.get("mpd_predict_sweep_set_sharded",{}).items()})
except Exception as e:
    print("parse failed", e)
PY
