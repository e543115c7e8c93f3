#!/bin/bash
# round 2, call K: TIM unrolled interior/edge walk: parity + timing + ncu
mkdir -p gpurun_out
echo "== tim tests"; timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "tim" --timeout 900 -p no:cacheprovider > gpurun_out/pytest_tim.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_tim.log
echo "== kernel table"; timeout 600 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -E "dwconv2d_sep k=15 .(ring|unrolled band walk, paired|register-sliding from global memory, FFMA2, factors as)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dwconv_sep_rg4" -c 1 -o gpurun_out/prof_tim_rg2 -f python tools/prof_fused.py tim2 > gpurun_out/ncu_tim.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_tim.log
python tools/ncu_summary.py gpurun_out/prof_tim_rg2.ncu-rep > gpurun_out/ncu_tim_rg2_summary.txt 2>&1; head -30 gpurun_out/ncu_tim_rg2_summary.txt
