#!/bin/bash
mkdir -p gpurun_out
echo "== kernels"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 240 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_kernels.log
echo "== kernel table"; timeout 300 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; tail -18 gpurun_out/kernels.log
echo "== ncu tim"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dwconv_sep" -s 2 -c 1 -o gpurun_out/prof_tim_r1 python tools/prof_fused.py tim > gpurun_out/ncu_tim.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_tim.log
echo "== ncu dim"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dim_" -s 4 -c 2 -o gpurun_out/prof_dim_r1 python tools/prof_fused.py dim > gpurun_out/ncu_dim.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_dim.log
