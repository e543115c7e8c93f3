#!/bin/bash
# round 2, call S: TIM default = the spill-free walk (loads one row ahead): parity of every launch path + timing, then the kernel + e2e suites
mkdir -p gpurun_out
echo "== tim tests"; timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "tim" --timeout 900 -p no:cacheprovider > gpurun_out/pytest_tim.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_tim.log
echo "== e2e (tim / ditimi / graph)"; timeout 1200 python -m pytest tests/test_e2e_gpu.py tests/test_e2e_baseline_gpu.py -m gpu -q -k "tim or ditimi or config3" --timeout 900 -p no:cacheprovider > gpurun_out/pytest_e2e_tim.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_e2e_tim.log
echo "== kernel table"; timeout 600 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -E "dwconv2d_sep k=15 .(unrolled|the same walk|register-sliding from global memory, FFMA2, factors as)"
