"""GPU diagnostic: which FMA-contraction variant of the bilinear blend reproduces torch's CUDA upsample kernel bit for bit?
Compares ta_dim_fwd (dim.blend = 0..4) with F.interpolate -> F.pad -> F.interpolate on the same device."""
import json, os, sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transferattack_b200 import ops, _lib

be = ops.backend()
res = {}
torch.manual_seed(0)
for S, rnd, R, top, left in [(224, 235, 246, 5, 6), (224, 224, 246, 0, 21), (224, 245, 246, 1, 0), (64, 67, 70, 1, 2), (299, 310, 328, 7, 9)]:
    x = torch.rand(4, 3, S, S, device="cuda")
    y1 = F.interpolate(x, size=[rnd, rnd], mode="bilinear", align_corners=False)
    y2 = F.pad(y1, [left, R - rnd - left, top, R - rnd - top], value=0)
    ref = F.interpolate(y2, size=[S, S], mode="bilinear", align_corners=False)
    for mode in range(5):
        _lib.tune_set("dim.blend", mode)
        out = be.dim(x, rnd, R, top, left, True)
        nd = int((out.view(torch.int32) != ref.view(torch.int32)).sum())
        res["S%d_rnd%d_mode%d" % (S, rnd, mode)] = {"n_diff_bits": nd, "max_abs": float((out - ref).abs().max())}
    # adjoint vs autograd on CUDA
    xg = x.clone().requires_grad_(True)
    y = F.interpolate(F.pad(F.interpolate(xg, size=[rnd, rnd], mode="bilinear", align_corners=False), [left, R - rnd - left, top, R - rnd - top]), size=[S, S], mode="bilinear", align_corners=False)
    g = torch.randn_like(y)
    (gin_ref,) = torch.autograd.grad(y, xg, g)
    gin = be.dim(g, rnd, R, top, left, False)
    res["S%d_rnd%d_bwd" % (S, rnd)] = {"max_abs": float((gin - gin_ref).abs().max()), "ref_absmax": float(gin_ref.abs().max())}
_lib.tune_set("dim.blend", 0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/diag_dim.json", "w"), indent=1)
for k, v in res.items():
    print(k, v)
