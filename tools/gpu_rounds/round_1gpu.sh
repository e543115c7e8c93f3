#!/bin/bash
mkdir -p gpurun_out
echo "== kernels"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 240 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_kernels.log
echo "== kernel table"; timeout 300 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -v "fused stream variant" gpurun_out/kernels.log | tail -20
echo "== sweep";   timeout 600 python bench.py --sweep > gpurun_out/sweep.log 2>&1; echo "rc=$?"; tail -11 gpurun_out/sweep.log
echo "== bench";   timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-2500
