"""Ensemble MI-FGSM with one surrogate per GPU (BASELINE config 5): images/s of
  (a) nccl  — ShardedEnsembleModel: NCCL all-reduce of logits (fwd) and of the input gradient (bwd), replicated fused update;
  (b) p2p   — FusedP2PEnsembleLoop: ta_fused_allreduce_update_linf (reduce-scatter + update + all-gather in one kernel
              over NVLink peer memory), logits by all_gather;
  (c) single — the reference's layout: all K members sequentially on ONE GPU (rank 0 only), same kernels.
Launch: python -m torch.distributed.run --nproc-per-node K tools/bench_ens.py [--batch 64] [--steps 5]
Writes gpurun_out/bench_ens_K.json (rank 0)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import transferattack_b200 as tab
from transferattack_b200 import multigpu

MODELS = ["resnet50", "resnet152", "inception_v3", "vit_b_16", "resnet18", "mobilenet_v2", "vgg16", "resnet101"]


def net(arch, seed, dev):
    torch.manual_seed(seed)
    kw = {"aux_logits": True, "init_weights": False} if arch == "inception_v3" else {}
    return getattr(torchvision.models, arch)(weights=None, **kw).eval().to(dev)


def timed(fn, steps, dev):
    dist.barrier(); torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record(); torch.cuda.synchronize(dev)
    ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--epoch", type=int, default=10)
    args = ap.parse_args()
    rank, local, K = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(args.batch, 3, 224, 224, generator=g).to(dev)
    y = torch.randint(0, 1000, (args.batch,), generator=g).to(dev)
    member = tab.utils.wrap_model(net(MODELS[rank], rank, dev))
    ens_cls = tab.load_attack_class("ens")
    out = {"K": K, "batch": args.batch, "epoch": args.epoch, "members": MODELS[:K]}

    a_nccl = multigpu.make_ens_attack(ens_cls, member, epoch=args.epoch)
    a_nccl.mean_mode = "exact"
    d_nccl = a_nccl(x, y); a_nccl(x, y)
    ms = timed(lambda: a_nccl(x, y), args.steps, dev)
    out["nccl"] = {"ms_per_attack": ms, "images_per_s": args.batch / ms * 1e3}

    a_p2p = multigpu.make_fused_p2p_ens(ens_cls, member, epoch=args.epoch)
    d_p2p = a_p2p(x, y); a_p2p(x, y)
    ms = timed(lambda: a_p2p(x, y), args.steps, dev)
    out["p2p"] = {"ms_per_attack": ms, "images_per_s": args.batch / ms * 1e3}
    out["p2p_vs_nccl_mismatch"] = int((d_p2p != d_nccl).sum())

    # the exchange + update step in isolation (same gradient tensor on every rank; median of 20 after 5 warm-ups)
    from transferattack_b200 import _lib, ops
    be = ops.backend(); lib = _lib.load()
    gfull = torch.randn_like(x) * 1e-4
    m = torch.zeros_like(x); d = torch.zeros_like(x); xa = torch.empty_like(x); so = torch.empty(args.batch, device=dev)
    st = a_p2p._buffers(x)
    lo, hi = multigpu.shard_bounds(args.batch, rank, K)
    n = x[0].numel()
    stream = torch.cuda.current_stream(dev)

    def step_nccl():
        gg = gfull.clone()
        dist.all_reduce(gg)
        be.fused_update_linf(gg, m, m, d, d, x, xa, None, so, 1.0, 1.6 / 255, 16 / 255, 0, 1.0)

    def step_p2p():
        st["G"].copy_(gfull)
        st["hg"].barrier(channel=0)
        _lib.check(lib.ta_fused_allreduce_update_linf(st["g_ptrs"], st["x_ptrs"], K, m.data_ptr(), m.data_ptr(), d.data_ptr(), d.data_ptr(),
                                                      x.data_ptr(), None, so.data_ptr(), 0, 1.0, 1.6 / 255, 16 / 255, 0.0, 1.0, lo, hi - lo, n,
                                                      stream.cuda_stream), "p2p")
        st["hx"].barrier(channel=0)

    for name, fn in (("nccl_allreduce_plus_update_us", step_nccl), ("p2p_fused_us", step_p2p)):
        for _ in range(5):
            fn()
        ts = []
        for _ in range(20):
            dist.barrier(); torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(dev)
            t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()) * 1e3)
        out[name] = sorted(ts)[len(ts) // 2]

    # the reference's layout: every member on one device, sequentially (rank 0 measures, the others idle)
    if rank == 0:
        nets = [tab.utils.wrap_model(net(MODELS[k], k, dev)) for k in range(K)]
        P = type("SingleENS", (ens_cls,), {"load_model": lambda self, _n: tab.utils.EnsembleModel(nets)})
        a_one = P(model_name="all-on-one", epoch=args.epoch)
        a_one.mean_mode = "exact"
        d_one = a_one(x, y); a_one(x, y)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            a_one(x, y)
        e1.record(); torch.cuda.synchronize(dev)
        ms1 = e0.elapsed_time(e1) / args.steps
        out["single_gpu_all_members"] = {"ms_per_attack": ms1, "images_per_s": args.batch / ms1 * 1e3}
        out["p2p_vs_single_mismatch"] = int((d_p2p != d_one).sum())
        out["nccl_vs_single_mismatch"] = int((d_nccl != d_one).sum())
    dist.barrier()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/bench_ens_%d.json" % K, "w"), indent=1)
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
