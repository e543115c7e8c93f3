#!/bin/bash
mkdir -p gpurun_out
echo "== bench 4 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n4.log | cut -c1-400
echo "== bench ens K=4"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 tools/bench_ens.py --batch 64 --steps 3 > gpurun_out/bench_ens4.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ens4.log | cut -c1-1500
