#!/bin/bash
# round 2, call P: Normalize's adjoint leaves the column sums of |g| (ta_normalize_bwd_colsum) -> the tail finishes the mean from them
mkdir -p gpurun_out
echo "== kernel tests"; timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "column_sums or abs_mean" --timeout 900 -p no:cacheprovider > gpurun_out/pytest_colsum.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_colsum.log
echo "== e2e"; timeout 1800 python -m pytest tests/test_e2e_gpu.py tests/test_e2e_baseline_gpu.py -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/pytest_e2e.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_e2e.log
echo "== bench"; timeout 1800 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['roofline'].get('traffic'), d['parity']['bit_identical'], d['graph']['captured'])
print(json.dumps(d.get('alt_mean_modes'))[:1800])
PY
tail -3 gpurun_out/bench.err
echo "== kernel table"; timeout 900 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -E "normalize_bwd|colsums|abs_mean|stream, scale"
echo "== ncu"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"normalize_bwd_colsum|aten_colsum_tree|ew_rows_kernel" -c 6 -o gpurun_out/prof_colsum_r2 -f python tools/prof_fused.py colsum > gpurun_out/ncu_colsum.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_colsum.log
python tools/ncu_summary.py gpurun_out/prof_colsum_r2.ncu-rep > gpurun_out/ncu_colsum_summary.txt 2>&1; grep -E "^====|gpu__time_duration|dram__bytes|smsp__inst_executed|issue_active|registers" gpurun_out/ncu_colsum_summary.txt
