"""GPU diagnostic: is oracle/aten_reduce.py's restatement of ATen's CUDA mean reduction order bit-equal to torch itself?

For every (B, n) and every variant of the order-relevant choices (shuffle direction, x-before-y, multiply-by-factor vs divide;
the default = what the installed torch's Reduce.cuh does) it compares
``emulate(|g|)`` with ``g.abs().mean(dim=(1,2,3))`` (resp. ``.mean(dim=1)``) bitwise and writes gpurun_out/diag_aten_mean.json.
Run on the GPU box: python tools/diag_aten_mean.py
"""
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aten_reduce  # noqa: E402


def main():
    dev = torch.device("cuda")
    prop = torch.cuda.get_device_properties(dev)
    sm, mt = prop.multi_processor_count, prop.max_threads_per_multi_processor
    out = {"torch": torch.__version__, "device": prop.name, "sm_count": sm, "max_threads_per_sm": mt, "cases": []}
    shapes = [(B, (3, 224, 224)) for B in (1, 2, 3, 4, 8, 16, 31, 32, 64, 65, 128, 256, 512, 600, 1024)]
    shapes += [(B, (3, 299, 299)) for B in (4, 64)] + [(B, (3, 32, 32)) for B in (4, 64)] + [(64, (3, 64, 64)), (16, (1, 224, 224)), (8, (3, 384, 384))]
    variants = [dict(shfl_descending=a, x_first=y, mul_factor=m) for a, y, m in itertools.product((True, False), (True, False), (True, False))]
    for B, chw in shapes:
        n = chw[0] * chw[1] * chw[2]
        cfg = aten_reduce.config(B, n, sm, mt)
        rec = {"B": B, "chw": chw, "n": n, "config": cfg, "match": {}}
        if cfg is not None:
            for seed, scale in ((0, 1.0), (1, 1e-4), (2, 37.0)):
                g = torch.randn(B, *chw, device=dev, generator=torch.Generator(dev).manual_seed(seed)) * scale
                ref = g.abs().mean(dim=(1, 2, 3))
                ref2 = g.abs().reshape(B, n).mean(dim=1)
                rec.setdefault("mean_dim123_equals_mean_dim1", True)
                rec["mean_dim123_equals_mean_dim1"] &= bool(torch.equal(ref, ref2))
                for v in variants:
                    key = "desc%d_xfirst%d_mul%d" % (v["shfl_descending"], v["x_first"], v["mul_factor"])
                    em = aten_reduce.emulate(g.abs().reshape(B, n), sm, mt, **v)
                    ok = bool(torch.equal(em, ref))
                    rec["match"][key] = rec["match"].get(key, True) and ok
                # a non-contiguous (channels_last) gradient, for the record: ATen reduces it in memory order
            rec["default_variant_matches"] = rec["match"]["desc1_xfirst1_mul1"]
        out["cases"].append(rec)
        print(B, chw, cfg, {k: v for k, v in rec["match"].items() if v})
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/diag_aten_mean.json", "w"), indent=1)
    good = [c for c in out["cases"] if c["config"] is not None]
    print("default variant matches on %d of %d supported shapes" % (sum(c["default_variant_matches"] for c in good), len(good)))


if __name__ == "__main__":
    main()
