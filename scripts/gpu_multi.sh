#!/bin/bash
# 2-GPU round trip: peer-exchange + sharded parity check, then bench at N=2.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "multi_check exit $?" >> gpurun_out/multi_check.log
tail -25 gpurun_out/multi_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench2 exit $?" >> gpurun_out/bench_n2.log
tail -3 gpurun_out/bench_n2.log; tail -15 gpurun_out/bench_n2.err
AAE_B200_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 2 --steps 200 --warmup 10 --no-extra > gpurun_out/bench_n2_nccl.log 2> gpurun_out/bench_n2_nccl.err; echo "bench2 nccl exit $?" >> gpurun_out/bench_n2_nccl.log
tail -2 gpurun_out/bench_n2_nccl.log; tail -5 gpurun_out/bench_n2_nccl.err
