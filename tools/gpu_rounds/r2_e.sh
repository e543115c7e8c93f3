#!/bin/bash
# round 2, call E: torch-order mean after the tree rewrite + split strategy; kernel table; bench line; then (last) the tcgen05 spectrum transform
mkdir -p gpurun_out
echo "== kernel tests (fused / mean)"; timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "fused or abs_mean or gra or adaea" --timeout 600 -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_kernels.log
echo "== e2e strict"; timeout 1200 python -m pytest tests/test_e2e_gpu.py tests/test_e2e_baseline_gpu.py -m gpu -q -k "strict or fold or config2 or config4 or gra or fast" --timeout 900 -p no:cacheprovider > gpurun_out/pytest_e2e.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_e2e.log
echo "== kernel table"; timeout 600 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -E "fused_tail|abs_mean|ATen"
echo "== bench"; timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','gpu_launches')}, d['e2e']['value'], d['roofline'], d['parity']['bit_identical'], d['graph'])
print(json.dumps(d.get('alt_mean_modes'))[:700])
print(json.dumps(d.get('other_configs'))[:1800])
print(json.dumps(d.get('fast_mode'))[:700])
PY
tail -3 gpurun_out/bench.err
echo "== ncu full: fused (split + cluster)"; timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fused_cluster_kernel|FusedStreamOpT|aten_abs_mean" -s 5 -c 5 -o gpurun_out/prof_fused_r2b python tools/prof_fused.py fused > gpurun_out/ncu_fused.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_fused.log
echo "== spectrum (tcgen05) tests, own process"; timeout 600 python -m pytest tests/test_zz_spectrum_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_spectrum.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_spectrum.log
