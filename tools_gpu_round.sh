#!/bin/bash
mkdir -p gpurun_out
echo "== kernels"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 240 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_kernels.log
echo "== e2e";     timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q --timeout 400 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_e2e.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_e2e.log
echo "== kernel table"; timeout 300 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; tail -16 gpurun_out/kernels.log
