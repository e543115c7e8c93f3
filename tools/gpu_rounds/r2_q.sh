#!/bin/bash
# round 2, call Q: tcgen05 isolation — the spectrum (SSM) tests alone in a fresh process on a fresh box, then health checks of the GPU
mkdir -p gpurun_out
echo "== before"; nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm,memory.used --format=csv,noheader
echo "== spectrum (tcgen05) tests"; TA_B200_TEST_TCGEN05=1 timeout 600 python -m pytest tests/test_zz_spectrum_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_spectrum.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_spectrum.log
echo "== health: new process, cuBLAS GEMM + our kernels + a second tcgen05 launch"; timeout 300 python - <<'PY'
import torch, time
a = torch.randn(4096, 4096, device="cuda"); b = torch.randn(4096, 4096, device="cuda")
torch.cuda.synchronize(); t = time.time(); c = (a @ b).sum().item(); torch.cuda.synchronize(); print("gemm ok", round(time.time() - t, 3), "s")
from transferattack_b200 import ops
be = ops.backend()
g = torch.randn(8, 3, 224, 224, device="cuda")
print("abs_mean ok", float(be.abs_mean(g).sum()) > 0)
x = torch.rand(2, 3, 224, 224, device="cuda")
y = be.spectrum_transform(x, torch.zeros_like(x), torch.ones_like(x))
torch.cuda.synchronize(); print("spectrum roundtrip max err", float((y - x).abs().max()))
PY
echo "rc=$?"
echo "== after"; nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm,memory.used --format=csv,noheader; nvidia-smi -q -d ERRORS 2>/dev/null | head -20 || true
dmesg 2>/dev/null | tail -5 || true
