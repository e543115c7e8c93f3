#!/bin/bash
# round 2, call G: mean-kernel load batching; kernel table; bench; fast-mode diagnostic
mkdir -p gpurun_out
echo "== kernel tests (fused / mean)"; timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "fused or abs_mean" --timeout 600 -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_kernels.log
echo "== DIM in the CUDA graph"; timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_kernels_gpu.py tests/test_e2e_baseline_gpu.py -m gpu -q -k "dim or graph or ditimi" --timeout 600 -p no:cacheprovider > gpurun_out/pytest_dim.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_dim.log
echo "== kernel table"; timeout 600 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -E "fused_tail|abs_mean|ATen"
echo "== bench"; timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','gpu_launches')}, d['e2e']['value'], d['roofline'], d['parity']['bit_identical'], d['graph'])
print(json.dumps(d.get('alt_mean_modes'))[:700])
PY
echo "== fast mode diag"; timeout 600 python tools/diag_fast_mode.py 2>&1 | tail -12
