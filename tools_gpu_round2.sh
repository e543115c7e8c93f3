#!/bin/bash
mkdir -p gpurun_out
echo "== kernel table (gpu0)"; CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; tail -32 gpurun_out/kernels.log
echo "== multigpu tests"; timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --timeout 600 --timeout-method thread -p no:cacheprovider -k "p2p" > gpurun_out/pytest_multigpu_p2p.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_multigpu_p2p.log
