#!/bin/bash
# round 2, 2-GPU run: sharded parity check (incl. set-sharded replicas), then bench at N=2 with parity_check
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  tests/multi_gpu_check.py > gpurun_out/r2m_multi_check.log 2>&1; echo "multi_check exit $?" >> gpurun_out/r2m_multi_check.log
tail -12 gpurun_out/r2m_multi_check.log | cut -c1-300
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; echo "bench2 exit $?"
tail -c 1500 gpurun_out/r2m_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2m_bench_n2.json') if l.startswith('{')][-1])
    print("N=2 MPD value %.0f e2e %.0f ms %.4f" % (d['value'], d['e2e']['value'], d['ms_per_step']))
    print("parity_check", d.get("parity_check"))
    for k in ("mpd_b1000","pubmed","pubmed_b500"):
        x=d.get(k)
        if x: print(k, "value %.0f ms %.3f" % (x['value'], x['ms_per_step']))
    print("sweep items", {k:(round(v['value']),v['path']) for k,v in d.get("mpd_predict_sweep",{}).items()})
    print("sweep sets ", {k:(round(v['value']),v['path']) for k,v in d.get("mpd_predict_sweep_set_sharded",{}).items()})
except Exception as e:
    print("parse failed", e)
PY
