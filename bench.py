#!/usr/bin/env python
"""bench.py — adversarial images / second, MI-FGSM ResNet-50 224^2 10 iterations (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (N>1: launched by torch.distributed.run)
    python bench.py --impl reference [...]                          the reference's CPU path (oracle port) on the host cores
    python bench.py --kernels                                       per-kernel roofline table  (gpurun_out/kernels.json)
    python bench.py --sweep                                         fused-update tuning sweep  (gpurun_out/sweep.json)

A "step" is one complete attack (10 iterations, each = surrogate forward + backward + the per-iteration kernels) on
one batch of B=64 synthetic 3x224x224 images per GPU (configs[1] of BASELINE.json). The batch shards across ranks with
no data-path collective (weak scaling: 64 images per GPU). One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_ELEMS = 3 * 224 * 224
FUSED_BYTES_PER_ELEM = 28           # reads g, m, delta, x (16) + writes m', delta', x_adv (12)  — DESIGN.md §kernels
METRIC = "adv images/sec, MI-FGSM ResNet-50 224^2 10-iter"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=64, help="images per GPU")
    p.add_argument("--arch", default="resnet50")
    p.add_argument("--attack", default="mifgsm")
    p.add_argument("--epoch", type=int, default=10)
    p.add_argument("--mean-mode", default="torch", choices=["torch", "exact"])
    p.add_argument("--graph", type=int, default=int(os.environ.get("TA_B200_GRAPH", "1")))
    p.add_argument("--kernels", action="store_true")
    p.add_argument("--sweep", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-worker", default="", help="internal: 'budget_seconds,steps,warmup,max_b' → JSON on stdout")
    p.add_argument("--no-eager-gpu", action="store_true")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------------------------
def peaks():
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        d = json.load(open(f))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = None
        self.p = None

    def __enter__(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        return self

    def __exit__(self, *a):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.f is None:
            return out
        try:
            self.f.flush()
            rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
            os.unlink(self.f.name)
        except Exception:
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def host_cores():
    """threads the CPU legs may really use: scheduler affinity, capped by the cgroup CPU quota"""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per))))
    except Exception:
        pass
    return n


def build_attack(pkg, name, net, **kw):
    cls = pkg.load_attack_class(name)
    wrap = pkg.utils.wrap_model
    # load_model is the documented override point (attack.py:40-65); the class supplying the surrogate declares it capturable
    P = type("Bench" + cls.__name__, (cls,), {"load_model": lambda self, _n: wrap(net), "graph_safe": True})
    return P(model_name="synthetic", **kw)


def make_net(arch, device):
    import torchvision
    torch.manual_seed(0)
    return getattr(torchvision.models, arch)(weights=None).eval().to(device)


def synth(B, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, 224, 224, generator=g), torch.randint(0, 1000, (B,), generator=g)


# ------------------------------------------------------------------------------------------------------------------
class EagerHooksMIFGSM:
    """The reference's eager-PyTorch hook chain on the GPU (attack.py:88,128,147-153 — 14 ATen launches per iteration),
    written out here as the comparator for the 'x vs the reference's PyTorch-GPU path' figure. Measured, never shipped."""

    def __init__(self, model, epsilon=16 / 255, alpha=1.6 / 255, epoch=10, decay=1.0):
        self.model, self.epsilon, self.alpha, self.epoch, self.decay = model, epsilon, alpha, epoch, decay
        self.loss = torch.nn.CrossEntropyLoss()

    def __call__(self, data, label):
        dev = next(self.model.parameters()).device
        data = data.clone().detach().to(dev); label = label.clone().detach().to(dev)
        delta = torch.zeros_like(data).requires_grad_(True)
        momentum = 0
        for _ in range(self.epoch):
            loss = self.loss(self.model(data + delta), label)
            grad = torch.autograd.grad(loss, delta)[0]
            momentum = momentum * self.decay + grad / (grad.abs().mean(dim=(1, 2, 3), keepdim=True))
            delta = torch.clamp(delta + self.alpha * momentum.sign(), -self.epsilon, self.epsilon)
            delta = torch.min(torch.max(delta, 0 - data), 1.0 - data).detach().requires_grad_(True)
        return delta.detach()


def cpu_reference_run(args, sample_b, steps, warmup):
    """oracle/torch_ref.py (eager restatement of the reference; the Python reference cannot travel to this box) on the
    host cores, all threads."""
    from oracle import torch_ref
    torch.set_num_threads(host_cores())
    net = make_net(args.arch, "cpu")
    atk = torch_ref.REF_ZOO[args.attack](torch_ref.ref_wrap_model(net), epoch=args.epoch)
    x, y = synth(sample_b)
    for _ in range(warmup):
        atk(x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        atk(x, y)
    dt = time.perf_counter() - t0
    return sample_b * steps / dt, dt / steps


def cpu_probe(args):
    """seconds per image for one full attack on the host, from a short probe (epoch=2 on 2 images)."""
    from oracle import torch_ref
    torch.set_num_threads(host_cores())
    net = make_net(args.arch, "cpu")
    atk = torch_ref.REF_ZOO[args.attack](torch_ref.ref_wrap_model(net), epoch=2)
    x, y = synth(4)
    atk(x, y)
    t0 = time.perf_counter(); atk(x, y); dt = time.perf_counter() - t0
    return dt / 4 / 2 * args.epoch


def cpu_worker(args):
    """child process: size a bounded sample from a probe, time it, print JSON"""
    budget, steps, warmup, max_b = [float(v) for v in args.cpu_worker.split(",")]
    steps, warmup, max_b = int(steps), int(warmup), int(max_b)
    per_img = cpu_probe(args)
    sample_b = int(max(1, min(max_b, budget / max(per_img * (steps + warmup), 1e-9))))
    val, per_step = cpu_reference_run(args, sample_b, steps, warmup)
    print(json.dumps({"value": val, "per_step_s": per_step, "sample_b": sample_b, "cores": host_cores(), "probe_s_per_img": per_img}), flush=True)


def cpu_leg(args, budget, steps, warmup, max_b, hard_timeout):
    """run the CPU leg in a child with a hard wall-clock bound (killed by PID on overrun)"""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", "%g,%d,%d,%d" % (budget, steps, warmup, max_b),
           "--arch", args.arch, "--attack", args.attack, "--epoch", str(args.epoch)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=hard_timeout, env=env)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (out.stderr or "no output")[-300:]}
    except subprocess.TimeoutExpired:
        return {"error": "cpu leg exceeded %ds" % hard_timeout}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    r = cpu_leg(args, 120.0, args.steps, args.warmup, args.batch, 420)
    if "error" in r:   # the oracle always exists; report the failure loudly but keep the line parseable
        r = {"value": float("nan"), "per_step_s": float("nan"), "sample_b": 0, "cores": host_cores(), "error": r["error"]}
    val, per_step, sample_b = r["value"], r["per_step_s"], r["sample_b"]
    cores = r["cores"]
    sample = "%d of %d images per step, %d iterations each" % (sample_b, args.batch, args.epoch)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MI-FGSM, ResNet-50, batch 64, 10 iters, eps=16/255 (BASELINE configs[1])", "attack": args.attack,
                   "arch": args.arch, "batch_per_gpu": args.batch, "epoch": args.epoch, "device": "cpu"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def timed_steps(fn, steps, dist, device):
    """barrier + synchronize, CUDA events around exactly `steps` calls on the launching stream, max over ranks."""
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    return float(ms.item())


def run_ours(args, rank, local_rank, world, dist):
    import transferattack_b200 as tab
    from transferattack_b200 import _lib
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    hbm_peak, peak_src = peaks()
    B = args.batch
    net = make_net(args.arch, device)
    atk = build_attack(tab, args.attack, net, epoch=args.epoch)
    atk.mean_mode = args.mean_mode
    atk.use_cuda_graph = bool(args.graph)
    x_host, y_host = synth(B, seed=1 + rank)
    x_pin, y_pin = x_host.pin_memory(), y_host.pin_memory()
    x_dev, y_dev = x_host.to(device), y_host.to(device)
    out_pin = torch.empty_like(x_pin).pin_memory()
    n_elem = B * IMG_ELEMS

    # -- warm-up (also triggers cuDNN heuristics / graph capture) ---------------------------------------------
    for _ in range(max(args.warmup, 3)):
        atk(x_dev, y_dev)
    torch.cuda.synchronize(device)

    # -- value: inputs resident in HBM ---------------------------------------------------------------------------
    launches0 = _lib.launch_count()
    with ClockSampler(local_rank) as clk:
        ms = timed_steps(lambda: atk(x_dev, y_dev), args.steps, dist, device)
    launches = _lib.launch_count() - launches0
    clocks = clk.summary()
    value = world * B * args.steps / (ms / 1e3)
    if args.graph:        # kernels replayed from the captured graph are not host launches: add the graph's own count per replay
        sts = list(getattr(atk, "_graphs", {}).values())
        if sts:
            launches += sts[-1].get("kernels_per_replay", 0) * args.epoch * args.steps

    # roofline of the dominant kernel: CUDA events around every ta_fused_update_linf launch, on its stream, live inside a
    # run of the same attack (eager launches of the same kernels: a graph replay cannot host per-launch events)
    kernel_events = []
    atk._kernel_events = kernel_events          # the base loop brackets its fused launch with CUDA events when set
    ms_ev = timed_steps(lambda: atk(x_dev, y_dev), max(2, args.steps // 2), dist, device)
    atk._kernel_events = None
    k_ms = [a.elapsed_time(b) for a, b in kernel_events]

    # -- e2e: host buffers through the plugin call, H2D of the batch and D2H of the perturbation inside the timed region
    def e2e_step():
        d = atk(x_pin, y_pin)
        out_pin.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(2):
        e2e_step()
    ms_e2e = timed_steps(e2e_step, args.steps, dist, device)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    roof = None
    if k_ms:
        avg_ms = float(np.mean(k_ms))
        achieved = FUSED_BYTES_PER_ELEM * n_elem / (avg_ms * 1e-3) / 1e9
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            traffic = json.load(open(tf)).get("fused_update_%s_B%d" % (args.mean_mode, B))
        roof = {"bound": "hbm", "kernel": "ta_fused_update_linf (%s)" % ("streaming, scale from torch" if args.mean_mode == "torch"
                                                                       else "cluster, in-kernel mean|g|"),
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "peak_source": peak_src, "avg_launch_us": avg_ms * 1e3, "launches_timed": len(k_ms),
                "algorithmic_bytes_per_launch": FUSED_BYTES_PER_ELEM * n_elem,
                "share_of_step": float(np.sum(k_ms)) / ms_ev}

    extra = {}
    if rank == 0 and world == 1:
        # the other mean mode, for the record (same timed procedure)
        other = "exact" if args.mean_mode == "torch" else "torch"
        atk.mean_mode = other
        for _ in range(2):
            atk(x_dev, y_dev)
        ms2 = timed_steps(lambda: atk(x_dev, y_dev), args.steps, None, device)
        ev2 = []
        atk._kernel_events = ev2
        timed_steps(lambda: atk(x_dev, y_dev), max(2, args.steps // 2), None, device)
        atk._kernel_events = None
        k2 = [a.elapsed_time(b) for a, b in ev2]
        ach2 = FUSED_BYTES_PER_ELEM * n_elem / (float(np.mean(k2)) * 1e-3) / 1e9 if k2 else None
        extra["alt_mean_mode"] = {"mode": other, "value": B * args.steps / (ms2 / 1e3), "ms_per_step": ms2 / args.steps,
                                  "roofline": {"achieved": ach2, "frac": ach2 / hbm_peak if ach2 else None,
                                               "avg_launch_us": float(np.mean(k2)) * 1e3 if k2 else None}}
        atk.mean_mode = args.mean_mode
        if not args.no_eager_gpu:
            # the reference's eager hook chain on this GPU (same surrogate, torchvision normalise incl. its host sync)
            from oracle import torch_ref  # comparator only
            eager = EagerHooksMIFGSM(torch_ref.ref_wrap_model(net), epoch=args.epoch)
            for _ in range(2):
                eager(x_dev, y_dev)
            ms3 = timed_steps(lambda: eager(x_dev, y_dev), args.steps, None, device)

            def eager_e2e():
                d = eager(x_host, y_host)
                d.cpu()
            eager_e2e()
            ms4 = timed_steps(eager_e2e, max(2, args.steps // 2), None, device)
            extra["reference_gpu_eager"] = {"value": B * args.steps / (ms3 / 1e3), "ms_per_step": ms3 / args.steps,
                                            "e2e_value": B * max(2, args.steps // 2) / (ms4 / 1e3),
                                            "note": "eager PyTorch hooks of attack.py on the same GPU/surrogate; informational"}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_leg(args, 20.0, 1, 0, 16, 240)
        if "error" in r:
            cpu_base = {"value": None, "unit": "images/s", "cores": host_cores(), "kind": "port", "sample": "failed: " + r["error"]}
        else:
            cpu_base = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                        "sample": "1 step of %d images (of %d), %d iterations, oracle/torch_ref.py on %d host threads"
                                  % (r["sample_b"], B, args.epoch, r["cores"])}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": ("MI-FGSM, ResNet-50, batch 64, 10 iters, eps=16/255 (BASELINE configs[1])"
                                    if (args.attack, args.arch, B, args.epoch) == ("mifgsm", "resnet50", 64, 10)
                                    else "%s, %s, batch %d, %d iters, eps=16/255" % (args.attack, args.arch, B, args.epoch)),
                       "attack": args.attack, "normalize_folded": bool(getattr(atk, "fold_normalize", False) and atk._fold_plan(x_dev) is not None),
                       "arch": args.arch, "batch_per_gpu": B, "global_batch": B * world, "epoch": args.epoch,
                       "parallelism": "batch-sharded x%d, no collective" % world, "mean_mode": args.mean_mode,
                       "surrogate": "torch autograd, fp32 (cuDNN TF32 convs as torch defaults), random-init weights",
                       "l2": "working set per step (activations of %d images) exceeds the 126 MB L2; no explicit flush" % B,
                       "cuda_graph": bool(args.graph)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * IMG_ELEMS * 4 + B * 8,
                    "d2h_bytes_per_step": B * IMG_ELEMS * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu_base,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def flush_l2(buf):
    # displace L2 with CLEAN lines (a read-only pass over 256 MB): a memset would leave 126 MB of dirty lines whose
    # write-back then competes with the timed kernel for DRAM
    buf.sum()


def time_kernel(fn, iters=20, flush=None):
    """per-launch duration with CUDA events on the launching stream; L2 flushed (256 MB memset) before each launch."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush_l2(flush)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def run_kernels(args):
    """roofline table of every kernel at BASELINE sizes (B=64): algorithmic bytes / measured time vs measured HBM peak."""
    from transferattack_b200 import ops, _lib
    import transferattack_b200.input_transformation.tim as tim
    be = ops.backend()
    hbm_peak, peak_src = peaks()
    B = args.batch
    N = B * IMG_ELEMS
    dev = "cuda"
    flush = torch.empty(1024 * 1024 * 1024 // 4, device=dev)     # 1 GB read-only pass (~170 us): displaces L2 with clean lines AND
    # keeps the GPU busy long enough that the host-side launch path of the next call is off the measured interval
    g = torch.randn(B, 3, 224, 224, device=dev) * 1e-4
    m = torch.randn_like(g); x = torch.rand_like(g); d = (torch.rand_like(g) * 2 - 1) * (16 / 255)
    m2, d2, xa = torch.empty_like(g), torch.empty_like(g), torch.empty_like(g)
    so = torch.empty(B, device=dev)
    scale = be.abs_mean(g)
    k2d, kcol, krow = tim.make_kernel("gaussian", 15)
    kc3 = torch.from_numpy(np.stack([kcol] * 3)).to(dev); kr3 = torch.from_numpy(np.stack([krow] * 3)).to(dev)
    k3 = torch.from_numpy(k2d.reshape(3, 15, 15)).to(dev)
    g5 = torch.randn(5 * B, 3, 224, 224, device=dev)
    a, al = 1.6 / 255, 16 / 255
    rows = []

    def add(name, bpe, fn, elems=N):
        med, mn = time_kernel(fn, flush=flush)
        gbs = bpe * elems / (med * 1e-3) / 1e9
        rows.append({"kernel": name, "bytes_per_elem": bpe, "elems": elems, "median_us": med * 1e3, "min_us": mn * 1e3,
                     "achieved_GBps": gbs, "frac_of_peak": gbs / hbm_peak})

    add("ATen reference: torch.add(x, d, out=) (same harness)", 12, lambda: torch.add(x, d, out=xa))
    add("ATen reference: tensor.copy_ (same harness)", 8, lambda: xa.copy_(x))
    add("fused_update_linf[cluster, in-kernel mean]", 28, lambda: be.fused_update_linf(g, m, m2, d, d2, x, xa, None, so, 1.0, a, al, 0, 1.0))
    add("fused_update_linf[stream, scale given]", 28, lambda: be.fused_update_linf(g, m, m2, d, d2, x, xa, scale, None, 1.0, a, al, 0, 1.0))
    for cap in (0, 4, 8, 16):
        for un in (1, 2, 4):
            _lib.tune_set("stream.cap", cap); _lib.tune_set("stream.unroll", un)
            add("  fused stream variant cap=%d unroll=%d" % (cap, un), 28,
                lambda: be.fused_update_linf(g, m, m2, d, d2, x, xa, scale, None, 1.0, a, al, 0, 1.0))
    _lib.tune_set("stream.cap", 0); _lib.tune_set("stream.unroll", 1)
    add("abs_mean_per_sample", 4, lambda: be.abs_mean(g))
    add("momentum", 12, lambda: be.momentum(g, m, scale, 1.0, out=m2))
    add("update_linf", 16, lambda: be.update_linf(d, x, m, a, al, 0, 1.0, out=d2))
    add("stage_add", 12, lambda: be.stage_add(x, d, out=xa))
    add("sim_fwd S=5", 24, lambda: be.sim(x, 5, True))
    add("sim_bwd S=5", 24, lambda: be.sim(g5, 5, False))
    for impl, bwd, fwdtab, tag in ((1, 0, 0, "direct; fwd tables = kernel parameters; adjoint = gather + scatter, tables in workspace"),
                                   (1, 1, 1, "direct; fwd tables in workspace; adjoint = independent gather"), (0, 0, 0, "4-pass")):
        _lib.tune_set("dim.impl", impl); _lib.tune_set("dim.bwd", bwd); _lib.tune_set("dim.fwdtab", fwdtab)
        add("dim_fwd [%s]" % tag, 8, lambda: be.dim(x, 235, 246, 5, 6, True))
        add("dim_bwd [%s]" % tag, 8, lambda: be.dim(g, 235, 246, 5, 6, False))
    _lib.tune_set("dim.impl", 1); _lib.tune_set("dim.bwd", 0); _lib.tune_set("dim.fwdtab", 0)
    hc, hr = kc3.cpu().numpy(), kr3.cpu().numpy()
    for band, f2, tag in ((3, 1, "register-sliding from global memory, FFMA2"), (3, 0, "register-sliding from global memory, FFMA"),
                          (2, 0, "register-sliding from TMA-staged smem")):
        _lib.tune_set("tim.band", band); _lib.tune_set("tim.f2", f2)
        add("dwconv2d_sep k=15 [%s, factors as kernel parameters]" % tag, 8, lambda: be.dwconv2d_sep(g, kc3, kr3, host=(hc, hr)))
        _lib.tune_set("tim.bh", 56)
        add("dwconv2d_sep k=15 [%s, parameters, band 56]" % tag, 8, lambda: be.dwconv2d_sep(g, kc3, kr3, host=(hc, hr)))
        _lib.tune_set("tim.bh", 32)
        add("dwconv2d_sep k=15 [%s, factors from device arrays]" % tag, 8, lambda: be.dwconv2d_sep(g, kc3, kr3))
    _lib.tune_set("tim.band", 3); _lib.tune_set("tim.f2", 1)
    for pf in (1, 0):
        _lib.tune_set("tim.prefetch", pf)
        add("dwconv2d_sep k=15 [from global memory, FFMA2, parameters, prefetch mode %d]" % pf, 8, lambda: be.dwconv2d_sep(g, kc3, kr3, host=(hc, hr)))
    _lib.tune_set("tim.prefetch", 2)
    _lib.tune_set("tim.band", 1)
    add("dwconv2d_sep k=15 [two-pass band kernel]", 8, lambda: be.dwconv2d_sep(g, kc3, kr3))
    _lib.tune_set("tim.band", 3); _lib.tune_set("tim.f2", 1)
    add("dwconv2d k=15 (direct)", 8, lambda: be.dwconv2d(g, k3))
    add("accumulate", 12, lambda: be.accumulate(m2, g, False))
    add("quantize_u8", 9, lambda: be.quantize_u8(x, d, True))
    # torch eager equivalents of the fused tail, for the same tensors
    def eager_tail():
        mm = m * 1.0 + g / g.abs().mean(dim=(1, 2, 3), keepdim=True)
        dd = torch.clamp(d + a * mm.sign(), -al, al)
        dd = torch.min(torch.max(dd, 0 - x), 1.0 - x)
        return x + dd
    med, mn = time_kernel(eager_tail, flush=flush)
    rows.append({"kernel": "torch eager tail (14 ATen launches, attack.py:88,128,147-153)", "bytes_per_elem": 128, "elems": N,
                 "median_us": med * 1e3, "min_us": mn * 1e3, "achieved_GBps": 128 * N / (med * 1e-3) / 1e9, "frac_of_peak": None})
    out = {"hbm_peak_GBps": hbm_peak, "peak_source": peak_src, "batch": B, "l2": "read-only pass over a 1 GB buffer before every timed launch (clean L2, host launch path hidden)", "rows": rows}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kernels.json"), "w"), indent=1)
    for r in rows:
        print("%-70s %8.1f us  %8.1f GB/s  %s" % (r["kernel"], r["median_us"], r["achieved_GBps"],
                                                   "" if r["frac_of_peak"] is None else "%.2f of peak" % r["frac_of_peak"]))


def run_sweep(args):
    from transferattack_b200 import ops, _lib
    be = ops.backend()
    hbm_peak, _ = peaks()
    dev = "cuda"
    flush = torch.empty(1024 * 1024 * 1024 // 4, device=dev)
    res = []
    for B in (64, 256):
        N = B * IMG_ELEMS
        g = torch.randn(B, 3, 224, 224, device=dev) * 1e-4
        m = torch.randn_like(g); x = torch.rand_like(g); d = (torch.rand_like(g) * 2 - 1) * (16 / 255)
        m2, d2, xa = torch.empty_like(g), torch.empty_like(g), torch.empty_like(g)
        so = torch.empty(B, device=dev)
        for variant in (0, 1):
            for cl in (2, 4, 8, 16):
                for threads, unroll in ((256, 1), (256, 2), (512, 1), (512, 2), (1024, 1)):
                    for k, v in (("fused.variant", variant), ("fused.cluster", cl), ("fused.threads", threads), ("fused.unroll", unroll)):
                        _lib.tune_set(k, v)
                    try:
                        med, mn = time_kernel(lambda: be.fused_update_linf(g, m, m2, d, d2, x, xa, None, so, 1.0, 1.6 / 255, 16 / 255, 0, 1.0),
                                              iters=10, flush=flush)
                    except RuntimeError as e:
                        res.append({"B": B, "variant": variant, "cluster": cl, "threads": threads, "unroll": unroll, "error": str(e)[:120]})
                        continue
                    gbs = 28 * N / (med * 1e-3) / 1e9
                    res.append({"B": B, "variant": variant, "cluster": cl, "threads": threads, "unroll": unroll,
                                "median_us": med * 1e3, "GBps": gbs, "frac": gbs / hbm_peak})
        del g, m, x, d, m2, d2, xa
    for k, v in (("fused.variant", 0), ("fused.cluster", 0), ("fused.threads", 512), ("fused.unroll", 2)):
        _lib.tune_set(k, v)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
    ok = [r for r in res if "GBps" in r]
    for B in (64, 256):
        best = sorted([r for r in ok if r["B"] == B], key=lambda r: -r["GBps"])[:5]
        for r in best:
            print(B, r)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.cpu_worker:
        cpu_worker(args)
        return
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.kernels:
        run_kernels(args); return
    if args.sweep:
        run_sweep(args); return
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world, dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
