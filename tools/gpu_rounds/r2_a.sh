#!/bin/bash
# round 2, call A: pin ATen's mean reduction order; parity at BASELINE's own configs with the round-1 kernels
mkdir -p gpurun_out
echo "== diag aten mean"; timeout 600 python tools/diag_aten_mean.py > gpurun_out/diag_aten_mean.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/diag_aten_mean.log
echo "== baseline-config parity"; timeout 1500 python -m pytest tests/test_e2e_baseline_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_baseline.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_baseline.log
