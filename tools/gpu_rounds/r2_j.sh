#!/bin/bash
# round 2, call J: ncu --set full on the separable-pass DIM kernels
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dim_fwd_sep|dim_bwd_sep" -c 2 -o gpurun_out/prof_dimsep_r2 -f python tools/prof_fused.py dim > gpurun_out/ncu_dimsep.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_dimsep.log
python tools/ncu_summary.py gpurun_out/prof_dimsep_r2.ncu-rep > gpurun_out/ncu_dimsep_summary.txt 2>&1; cat gpurun_out/ncu_dimsep_summary.txt | head -70
