#!/bin/bash
mkdir -p gpurun_out
echo "== diag dim"; timeout 300 python tools/diag_dim_aten.py > gpurun_out/diag_dim.log 2>&1; echo "rc=$?"; grep -E "mode|bwd" gpurun_out/diag_dim.log | head -40
echo "== bench";   timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log
echo "== sweep";   timeout 600 python bench.py --sweep > gpurun_out/sweep.log 2>&1; echo "rc=$?"; tail -16 gpurun_out/sweep.log
echo "== ncu full fused"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fused_cluster_kernel|FusedStreamOp" -s 4 -c 4 -o gpurun_out/prof_fused_r1 python tools/prof_fused.py fused > gpurun_out/ncu_fused.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_fused.log
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-gpu > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -c 300 gpurun_out/ncu_bench.log
