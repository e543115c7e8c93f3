#!/bin/bash
mkdir -p gpurun_out
echo "== kernels"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 240 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_kernels.log
echo "== e2e";     timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q --timeout 400 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_e2e.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_e2e.log
echo "== kernel table"; timeout 300 python bench.py --kernels > gpurun_out/kernels.log 2>&1; echo "rc=$?"; tail -18 gpurun_out/kernels.log
for cfg in "resnet50 64" "resnet50 32" "resnet50 8" "resnet18 32" "resnet18 8"; do set -- $cfg
  for g in 0 1; do echo "== bench $1 B=$2 graph=$g"; timeout 300 python bench.py --arch $1 --batch $2 --graph $g --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu > gpurun_out/bench_$1_$2_g$g.log 2>&1; echo "rc=$?"; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/bench_$1_$2_g$g.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("value %.1f img/s  ms/step %.2f  e2e %.1f  roof %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"] if d["roofline"] else -1))
except Exception as e: print("no line", e); print(open("gpurun_out/bench_$1_$2_g$g.log").read()[-600:])
PY
  done
done
