#!/bin/bash
# round 2, call H: mean kernel as 4-CTA clusters; kernel table; bench
mkdir -p gpurun_out
echo "== kernel tests"; timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "fused or abs_mean or dim_dyn" --timeout 600 -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_kernels.log
echo "== kernel table"; timeout 600 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; grep -E "us " gpurun_out/kernels.log | grep -E "fused_tail|abs_mean|ATen"
echo "== bench"; timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['parity']['bit_identical'], d['graph']['captured'])
print(json.dumps(d.get('alt_mean_modes'))[:900])
print(json.dumps(d.get('other_configs',{}).get('config3_ditimi_resnet50_b32_per_gpu_share'))[:300])
PY
