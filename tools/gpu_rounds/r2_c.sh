#!/bin/bash
# round 2, call C: full GPU suite, bench line (+ reference arm), kernel table, ncu launch list + full captures of the shipped kernels
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
echo "== kernel table"; timeout 600 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -v "variant cap"
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --graph 0 --no-cpu-baseline --no-eager-gpu --no-extras > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches_r2.csv
echo "== ncu full: shipped fused kernels"; timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fused_cluster_kernel|FusedStreamOpT|aten_abs_mean" -s 4 -c 4 -o gpurun_out/prof_fused_r2 python tools/prof_fused.py fused > gpurun_out/ncu_fused.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_fused.log
echo "== ncu full: tim dim"; timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"dwconv_sep_rg|dim_fwd_direct|dim_bwd" -s 8 -c 4 -o gpurun_out/prof_timdim_r2 python tools/prof_fused.py timdim > gpurun_out/ncu_timdim.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_timdim.log
