#!/bin/bash
# 2 GPUs: multi-GPU tests (NCCL sharding, NCCL ensemble, fused P2P ensemble) and the sharded bench after the session-2 changes
mkdir -p gpurun_out
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_multigpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_multigpu.log
echo "== bench 2 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-400
echo "== bench reference arm under torchrun"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-200
