#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus2.txt
echo "== multigpu tests"; timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --timeout 600 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_multigpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_multigpu.log
echo "== bench 2 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_n2.log | cut -c1-900
echo "== bench ref arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-300
echo "== bench 1 gpu"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-1500
