#!/bin/bash
# round 2, call M: the final tree — whole GPU suite, smoke, bench (both arms), kernel table, launch list, ncu of the hot kernels
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu_r2.txt 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu_r2.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench"; timeout 1800 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['roofline'].get('traffic'), d['parity']['bit_identical'], d['graph']['captured'], d.get('clocks'))
print(json.dumps(d.get('alt_mean_modes'))[:1200])
print(json.dumps(d.get('other_configs'))[:2500])
print(json.dumps(d.get('fast_mode'))[:900])
print(json.dumps(d.get('cpu_baseline'))[:400])
PY
tail -3 gpurun_out/bench.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -c 900 gpurun_out/bench_ref.json
echo "== kernel table"; timeout 900 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -vE "variant|prefetch mode|band 56|device arrays|FFMA,|TMA-staged"
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --graph 0 --no-cpu-baseline --no-eager-gpu --no-extras > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches_r2.csv
echo "== ncu full: tail kernels + TIM + DIM"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ew_rows_kernel|normalize_bwd_colsum|aten_abs_mean|dwconv_sep_rg2|dim_fwd_direct|dim_bwd_sep" -c 8 -o gpurun_out/prof_final_r2 -f python tools/prof_fused.py final > gpurun_out/ncu_final.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_final.log
python tools/ncu_summary.py gpurun_out/prof_final_r2.ncu-rep > gpurun_out/ncu_final_summary.txt 2>&1; grep -E "^====|gpu__time_duration|dram__bytes|smsp__inst_executed|issue_active" gpurun_out/ncu_final_summary.txt
