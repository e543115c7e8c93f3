#!/bin/bash
# round 2, call I: separable-pass DIM kernels: parity tests, kernel table rows
mkdir -p gpurun_out
echo "== dim tests"; timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "dim" --timeout 600 -p no:cacheprovider > gpurun_out/pytest_dim.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_dim.log
echo "== kernel table"; timeout 600 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -E "dim_|dwconv2d_sep k=15 \[register-sliding from global memory, FFMA2, factors as"
