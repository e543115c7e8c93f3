"""GPU diagnostic for the opt-in fast mode: forward + input-gradient time of ResNet-50 at B=64 in the candidate surrogate formats.
    python tools/diag_fast_mode.py  → gpurun_out/diag_fast_mode.json"""
import copy
import json
import os
import sys

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    torch.manual_seed(0)
    net = torchvision.models.resnet50(weights=None).eval().cuda()
    for p in net.parameters():
        p.requires_grad_(False)
    B = 64
    x = torch.rand(B, 3, 224, 224, device="cuda")
    y = torch.randint(0, 1000, (B,), device="cuda")
    ce = torch.nn.CrossEntropyLoss()
    out = {}

    def run(model, conv):
        def f():
            xx = x.clone().requires_grad_(True)
            loss = ce(model(conv(xx)).float(), y)
            torch.autograd.grad(loss, xx)
        return f

    out["fp32_nchw"] = timed(run(net, lambda t: t))
    n_cl = copy.deepcopy(net).to(memory_format=torch.channels_last)
    out["fp32_channels_last"] = timed(run(n_cl, lambda t: t.contiguous(memory_format=torch.channels_last)))
    n_bf = copy.deepcopy(net).to(torch.bfloat16)
    out["bf16_nchw"] = timed(run(n_bf, lambda t: t.to(torch.bfloat16)))
    n_bfcl = copy.deepcopy(net).to(torch.bfloat16).to(memory_format=torch.channels_last)
    out["bf16_channels_last"] = timed(run(n_bfcl, lambda t: t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)))
    n_hf = copy.deepcopy(net).to(torch.float16).to(memory_format=torch.channels_last)
    out["fp16_channels_last"] = timed(run(n_hf, lambda t: t.to(torch.float16).contiguous(memory_format=torch.channels_last)))

    def autocast_fn():
        xx = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = ce(n_cl(xx.contiguous(memory_format=torch.channels_last)).float(), y)
        torch.autograd.grad(loss, xx)
    out["autocast_bf16_channels_last"] = timed(autocast_fn)
    # CUDA-graphed bf16 channels_last (what the fast mode replays)
    g = torch.cuda.CUDAGraph()
    xs = x.clone().requires_grad_(True)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            loss = ce(n_bfcl(xs.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)).float(), y)
            torch.autograd.grad(loss, xs)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        loss = ce(n_bfcl(xs.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)).float(), y)
        gr = torch.autograd.grad(loss, xs)
    out["bf16_channels_last_graph"] = timed(lambda: g.replay())
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/diag_fast_mode.json", "w"), indent=1)


if __name__ == "__main__":
    main()
