#!/bin/bash
mkdir -p gpurun_out
echo "== smoke";   timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== kernel table"; timeout 300 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "dim_|dwconv|abs_mean" gpurun_out/kernels.log
echo "== bench";   timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-1800
