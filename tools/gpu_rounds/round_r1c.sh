#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests (tim, dim)"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "tim or dim" --timeout 300 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_kernels.log
echo "== kernel table"; timeout 400 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "dim_|dwconv|ATen" gpurun_out/kernels.log
