#!/bin/bash
# session 2 of round 1: new TIM (register-sliding) and DIM (direct) kernels — parity, then the per-kernel table
mkdir -p gpurun_out
echo "== kernel tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_kernels.log
echo "== e2e dim/tim"; timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q -k "dim or tim" --timeout 400 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_e2e_dim.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_e2e_dim.log
echo "== kernel table"; timeout 400 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "dim_|dwconv|abs_mean|ATen|quantize" gpurun_out/kernels.log
