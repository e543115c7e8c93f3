#!/bin/bash
# round 2, call B: ATen mean-order diagnostic, the new TA_MEAN_TORCH / ta_fused_tail kernels, e2e parity incl. BASELINE's own configs
mkdir -p gpurun_out
echo "== diag aten mean"; timeout 600 python tools/diag_aten_mean.py > gpurun_out/diag_aten_mean.log 2>&1; echo "rc=$?"; tail -32 gpurun_out/diag_aten_mean.log
echo "== kernel tests"; timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_kernels.log
echo "== e2e"; timeout 1500 python -m pytest tests/test_e2e_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_e2e.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_e2e.log
echo "== baseline-config parity"; timeout 1500 python -m pytest tests/test_e2e_baseline_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_baseline.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_baseline.log
