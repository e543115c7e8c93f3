#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "column_sums" --timeout 300 -p no:cacheprovider 2>&1 | tail -2
timeout 600 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -E "normalize_bwd"
