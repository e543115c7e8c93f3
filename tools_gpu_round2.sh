#!/bin/bash
mkdir -p gpurun_out
echo "== bench ens K=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/bench_ens.py --batch 64 --steps 5 > gpurun_out/bench_ens2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_ens2.log | cut -c1-1200
