#!/bin/bash
# round 1, session 2 — full GPU pass: every -m gpu test, smoke, the bench line (+ reference arm), kernel table, other configs,
# ncu launch list of the bench command
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench.json
echo "== kernel table"; timeout 400 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -v "variant cap"
echo "== other configs"
for cfg in "tim resnet50 64 10" "ditimi resnet50 32 10" "vmifgsm vit_b_16 32 2" "dim resnet50 64 10"; do set -- $cfg
  timeout 600 python bench.py --attack $1 --arch $2 --batch $3 --epoch $4 --steps 3 --warmup 2 --no-cpu-baseline --no-eager-gpu > gpurun_out/bench_cfg_$1.log 2>&1; echo "$1 rc=$?"; tail -1 gpurun_out/bench_cfg_$1.log | cut -c1-200
done
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 1 --graph 0 --no-cpu-baseline --no-eager-gpu > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches_r1b.csv
