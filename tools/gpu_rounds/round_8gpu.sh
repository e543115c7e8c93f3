#!/bin/bash
mkdir -p gpurun_out
echo "== bench 8 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n8.log | cut -c1-600
echo "== config 3: DI-TI-MI B=256 batch-sharded over 8 GPUs (32 per GPU)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --attack ditimi --batch 32 --steps 10 --warmup 3 --no-eager-gpu > gpurun_out/bench_cfg3_n8.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_cfg3_n8.log | cut -c1-400
