#!/bin/bash
# round 2, call L: DIM forward walk: parity + timing + ncu
mkdir -p gpurun_out
echo "== dim tests"; timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "dim" --timeout 900 -p no:cacheprovider > gpurun_out/pytest_dim.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_dim.log
echo "== kernel table"; timeout 600 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -E "dim_"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dim_fwd_walk" -c 1 -o gpurun_out/prof_dim_walk -f python tools/prof_fused.py dim > gpurun_out/ncu_dimw.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_dimw.log
python tools/ncu_summary.py gpurun_out/prof_dim_walk.ncu-rep > gpurun_out/ncu_dim_walk_summary.txt 2>&1; head -30 gpurun_out/ncu_dim_walk_summary.txt
