"""-m gpu: the whole attack through the plugin API (kernels via the C-ABI) against the eager-PyTorch restatement of the
reference (oracle/torch_ref.py) on the SAME device with the SAME surrogate, seeds and cuDNN settings.

strict mean mode ('torch'): the perturbation must be bit-identical (hence also after uint8 quantisation) for every
attack whose ops have a fully determined order; DIM/TIM (ATen's blend / conv order is not reproducible, and ATen's
bilinear backward uses atomics) are held to a mismatch fraction instead.  A JSON report goes to gpurun_out/."""
import json
import os

import numpy as np
import pytest
import torch
import torchvision

import transferattack_b200 as tab
from oracle import torch_ref
from conftest import ROOT
from helpers import make_attack, seed_all

pytestmark = pytest.mark.gpu
REPORT = {}


@pytest.fixture(scope="module", autouse=True)
def _setup():
    from transferattack_b200 import ops
    ops._install_backend_for_tests(None)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    yield
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "e2e_parity.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


def _net(arch="resnet18", seed=0):
    torch.manual_seed(seed)
    return getattr(torchvision.models, arch)(weights=None).eval().cuda()


def _data(B=4, S=224, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, S, S, generator=g), torch.randint(0, 1000, (B,), generator=g)


def _stats(d, dr, x):
    d, dr = d.float().cpu(), dr.float().cpu()
    diff = (d - dr).abs()
    q = torch_ref.save_images_u8(x, d); qr = torch_ref.save_images_u8(x, dr)
    return {"max_abs": float(diff.max()), "n_gt_1e-5": int((diff > 1e-5).sum()), "numel": d.numel(),
            "u8_mismatch": int((q != qr).sum()), "bit_identical": bool(torch.equal(d, dr))}


CASES = {
    "ifgsm": ("ifgsm", {}), "mifgsm": ("mifgsm", {}), "nifgsm": ("nifgsm", {}), "fgsm": ("fgsm", {}),
    "sim": ("sim", {"epoch": 4}), "admix": ("admix", {"epoch": 2}), "vmifgsm": ("vmifgsm", {"num_neighbor": 3, "epoch": 4}),
    "vnifgsm": ("vnifgsm", {"num_neighbor": 2, "epoch": 3}), "emifgsm": ("emifgsm", {"epoch": 4}),
    "mifgsm_rs": ("mifgsm", {"random_start": True}), "mifgsm_targeted": ("mifgsm", {"targeted": True}),
}


def _pair(key, net, mean_mode="torch", **over):
    name, kw = CASES[key]
    kw = dict(kw, **over)
    x, y = _data()
    lab = torch.stack([y, (y + 1) % 1000]) if kw.get("targeted") else y
    ref = torch_ref.REF_ZOO[name](torch_ref.ref_wrap_model(net), **kw)
    seed_all(2); torch.cuda.manual_seed_all(2)
    dr = ref(x, lab)
    atk = make_attack(tab, name, net, **kw)
    atk.mean_mode = mean_mode
    seed_all(2); torch.cuda.manual_seed_all(2)
    d = atk(x, lab)
    return d, dr, x


@pytest.mark.parametrize("mean_mode", ["torch", "aten"])
@pytest.mark.parametrize("key", sorted(CASES))
def test_strict_mode_is_bit_identical(key, mean_mode):
    """'torch': mean|g| formed inside the fused kernel in torch's own summation order (TA_MEAN_TORCH) — no ATen kernel in the
    tail; 'aten': the scale from torch's own op. Both must reproduce the reference bit for bit."""
    net = _net()
    d, dr, x = _pair(key, net, mean_mode=mean_mode)
    st = _stats(d, dr, x)
    REPORT["strict_%s/%s" % (mean_mode, key)] = st
    assert d.is_cuda and not d.requires_grad
    assert st["bit_identical"], st


def test_reference_noise_floor_and_exact_mean_mode():
    """How far the reference is from ITSELF run twice (must be 0 with deterministic cuDNN), and how far the fully fused
    'exact' mean mode lands from it (last-bit differences in mean|g| can flip the sign of near-zero momentum entries)."""
    net = _net()
    _, dr1, x = _pair("mifgsm", net)
    _, dr2, _ = _pair("mifgsm", net)
    REPORT["noise_floor/mifgsm_ref_vs_ref"] = _stats(dr1, dr2, x)
    assert torch.equal(dr1, dr2)
    d, dr, x = _pair("mifgsm", net, mean_mode="exact")
    st = _stats(d, dr, x)
    REPORT["exact/mifgsm"] = st
    assert st["max_abs"] <= 2 * 16 / 255 + 1e-6
    # measured floor on B200: 0 differing elements; the fp64-exact mean can differ from torch's fp32 tree in the last bit, which
    # only matters for momentum entries that are zero to rounding → at most a handful per million (was 2 % in round 1)
    assert st["n_gt_1e-5"] <= 1e-5 * st["numel"], st


def test_ens_two_members_bit_identical():
    nets = [_net("resnet18", 0), _net("mobilenet_v2", 3)]
    x, y = _data()
    ref = torch_ref.ref_mifgsm(torch_ref.RefEnsemble([torch_ref.ref_wrap_model(n) for n in nets]), epoch=4)
    dr = ref(x, y)
    atk = make_attack(tab, "ens", nets, epoch=4)
    d = atk(x, y)
    st = _stats(d, dr, x)
    REPORT["strict/ens"] = st
    assert st["bit_identical"], st


@pytest.mark.parametrize("name", ["dim", "tim", "ditimi"])
def test_dim_tim_single_iteration_tolerance(name):
    net = _net()
    x, y = _data()
    kw = dict(epoch=1)
    if name != "tim":
        kw["diversity_prob"] = 1.0
    ref = torch_ref.REF_ZOO[name](torch_ref.ref_wrap_model(net), **kw)
    seed_all(5); dr = ref(x, y)
    atk = make_attack(tab, name, net, **kw)
    seed_all(5); d = atk(x, y)
    st = _stats(d, dr, x)
    REPORT["one_iter/" + name] = st
    # forward is bit-identical to torch's CUDA kernels; only the adjoint's summation order differs (ATen: atomicAdd) →
    # sign flips of momentum entries that are zero to rounding
    assert st["n_gt_1e-5"] <= 2e-4 * st["numel"], st


@pytest.mark.parametrize("name", ["dim", "tim", "ditimi"])
def test_dim_tim_ten_iterations_report(name):
    net = _net()
    x, y = _data()
    kw = {} if name == "tim" else {"diversity_prob": 0.5}
    ref = torch_ref.REF_ZOO[name](torch_ref.ref_wrap_model(net), **kw)
    seed_all(6); dr = ref(x, y)
    seed_all(6); dr2 = ref(x, y)
    atk = make_attack(tab, name, net, **kw)
    seed_all(6); d = atk(x, y)
    REPORT["ten_iter/" + name] = _stats(d, dr, x)
    REPORT["ten_iter/" + name + "_ref_vs_ref"] = _stats(dr, dr2, x)
    assert float(d.abs().max()) <= 16 / 255 + 1e-7
    assert torch.isfinite(d).all()


def test_vit_b16_and_resnet50_run_strict():
    for arch, B in (("resnet50", 4), ("vit_b_16", 2)):
        net = _net(arch)
        x, y = _data(B)
        ref = torch_ref.ref_mifgsm(torch_ref.ref_wrap_model(net), epoch=3)
        dr = ref(x, y)
        d = make_attack(tab, "mifgsm", net, epoch=3)(x, y)
        st = _stats(d, dr, x)
        REPORT["strict/mifgsm_" + arch] = st
        assert st["bit_identical"], st


def test_host_input_device_output_like_reference_main():
    """main.py:52-53 passes CPU tensors and adds `perturbations.cpu()` to the CPU images."""
    net = _net()
    x, y = _data(2)
    d = make_attack(tab, "mifgsm", net, epoch=2)(x, y)
    assert d.is_cuda and d.shape == x.shape
    adv = x + d.cpu()
    assert float(adv.min()) >= 0 and float(adv.max()) <= 1.0 + 1e-6
    xp = x.pin_memory()
    d2 = make_attack(tab, "mifgsm", net, epoch=2)(xp, y)
    assert torch.equal(d, d2)


@pytest.mark.parametrize("name,kw", [("mifgsm", {}), ("nifgsm", {}), ("ifgsm", {}), ("tim", {"epoch": 4}), ("sim", {"epoch": 3}),
                                     ("mifgsm", {"random_start": True}), ("mifgsm", {"targeted": True})])
def test_cuda_graph_replay_is_bit_identical(name, kw):
    """use_cuda_graph replays one captured iteration `epoch` times: same kernels in the same order → same bits, also on
    the second batch through the cached graph, and against the reference restatement."""
    net = _net()
    x, y = _data()
    lab = torch.stack([y, (y + 1) % 1000]) if kw.get("targeted") else y
    plain = make_attack(tab, name, net, **kw)
    plain.use_cuda_graph = False
    seed_all(2); torch.cuda.manual_seed_all(2)
    d_plain = plain(x, lab)
    graphed = make_attack(tab, name, net, **kw)
    graphed.use_cuda_graph = True
    seed_all(2); torch.cuda.manual_seed_all(2)
    d_graph = graphed(x, lab)
    assert torch.equal(d_plain, d_graph)
    assert len(graphed._graphs) == 1
    x2, y2 = _data(seed=9)
    lab2 = torch.stack([y2, (y2 + 1) % 1000]) if kw.get("targeted") else y2
    seed_all(3); torch.cuda.manual_seed_all(3)
    d2_plain = plain(x2, lab2)
    seed_all(3); torch.cuda.manual_seed_all(3)
    d2_graph = graphed(x2, lab2)
    assert torch.equal(d2_plain, d2_graph) and len(graphed._graphs) == 1
    REPORT["graph/" + name + ("_" + "_".join(kw) if kw else "")] = {"bit_identical": True}


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
@pytest.mark.parametrize("mean_mode", ["torch", "torch-meankernel", "aten", "exact", "torch+adjoint"])
@pytest.mark.parametrize("name", ["mifgsm", "ifgsm", "tim"])
def test_normalize_fold_is_bit_identical(name, mean_mode, graph):
    """SURVEY §8 f1 on the GPU: fused tail emitting the normalised model input (+ Normalize's adjoint in 'exact' mode with
    the base get_grad) ≡ separate ta_normalize_* kernels, eager launches and CUDA-graph replay; strict mode ≡ the reference."""
    from transferattack_b200 import _lib
    net = _net()
    x, y = _data()
    res = {}
    for fold in (False, True):
        atk = make_attack(tab, name, net, epoch=4)
        if mean_mode == "torch+adjoint":                     # the in-kernel torch-order mean WITH Normalize's adjoint in the kernels
            atk.mean_mode = "torch"; atk.fold_adjoint = True
        elif mean_mode == "torch-meankernel":                # the separate torch-order mean kernel (no column sums from the adjoint)
            atk.mean_mode = "torch"; atk.colsum_adjoint = False
        else:
            atk.mean_mode = mean_mode
        atk.fold_normalize = fold; atk.use_cuda_graph = graph
        assert (atk._fold_plan(x.cuda()) is not None) == fold
        if fold:    # the default for 'torch' with the base get_grad: the adjoint kernel leaves the column sums (TIM smooths the gradient → no)
            assert bool(atk._fold_plan(x.cuda(), atk._mean_kernel_mode(x.cuda()))[5]) == (mean_mode == "torch" and name != "tim")
        before = _lib.launch_count()
        res[fold] = atk(x, y)
        res[fold, "launches"] = _lib.launch_count() - before
        assert bool(getattr(atk, "_graphs", None)) == graph
    assert torch.equal(res[False], res[True])
    if not graph:       # launches of OUR kernels per attack: the fold removes the Normalize forward (and adjoint when deferred)
        # Normalize's adjoint inside the tail kernels: default with the 'exact' cluster kernel, opt-in (fold_adjoint) with the torch-order mean
        deferred = mean_mode in ("exact", "torch+adjoint") and name != "tim"
        colsum = mean_mode == "torch" and name != "tim"        # the adjoint kernel finishes mean|g| itself: no mean kernel either
        assert res[False, "launches"] - res[True, "launches"] == 4 * (1 + int(deferred) + int(colsum)) - 1   # one extra Normalize forward up front
    if mean_mode in ("torch", "torch-meankernel", "aten", "torch+adjoint"):
        ref = torch_ref.REF_ZOO[name](torch_ref.ref_wrap_model(net), epoch=4)(x, y)
        assert torch.equal(res[True], ref)
    REPORT["fold/%s_%s_%s" % (name, mean_mode, "graph" if graph else "eager")] = {"bit_identical": True}


def test_pifgsm_native_matches_restatement_on_gpu():
    """PI-FGSM on the kernels against the eager restatement of gradient/pifgsm.py on the same GPU (the
    restatement is pinned to the live reference in tests/test_reference_live.py). Only the 3x3 projection convolution's
    summation order is free (cuDNN vs ours) and it only enters through sign(): a handful of elements per million at most."""
    net = _net()
    x, y = _data()
    for kw in ({}, {"decay": 1.0, "epoch": 4}):
        ref = torch_ref.RefPIFGSM(torch_ref.ref_wrap_model(net), **kw)(x, y)
        d = make_attack(tab, "pifgsm", net, **kw)(x, y)
        st = _stats(d, ref, x)
        REPORT["pifgsm" + ("_mpi" if kw else "")] = st
        assert int((d != ref).sum()) <= 1e-5 * d.numel(), st
        assert float(d.abs().max()) <= 16 / 255 + 1e-7


def test_gra_and_adaea_native_match_restatement_on_gpu():
    """SURVEY §8 f4 on the GPU: native GRA (ta_gra_update, in-kernel Philox neighbours) bit-identical to the restatement of
    gradient/gra.py; native AdaEA (ta_adaea_drf) equal to the restatement of ensemble/adaea.py up to pixels whose map value is
    within rounding of the threshold."""
    net = _net()
    x, y = _data()
    kw = {"num_neighbor": 3, "epoch": 4}
    seed_all(2); torch.cuda.manual_seed_all(2)
    ref = torch_ref.RefGRA(torch_ref.ref_wrap_model(net), **kw)(x, y)
    seed_all(2); torch.cuda.manual_seed_all(2)
    d = make_attack(tab, "gra", net, **kw)(x, y)
    REPORT["gra"] = _stats(d, ref, x)
    assert torch.equal(d, ref), REPORT["gra"]
    nets = [_net("resnet18", 0), _net("mobilenet_v2", 3), _net("resnet18", 5)]
    seed_all(3); torch.cuda.manual_seed_all(3)
    ref = torch_ref.RefAdaEA(torch_ref.RefEnsemble([torch_ref.ref_wrap_model(n) for n in nets]), epoch=3)(x, y)
    seed_all(3); torch.cuda.manual_seed_all(3)
    d = make_attack(tab, "adaea", nets, epoch=3)(x, y)
    REPORT["adaea"] = _stats(d, ref, x)
    assert REPORT["adaea"]["n_gt_1e-5"] <= 1e-5 * d.numel(), REPORT["adaea"]


def test_fast_mode_is_opt_in_and_keeps_the_attack_strength():
    """The opt-in bf16 / channels_last surrogate (Attack.fast_mode) is NOT a parity path: its acceptance is that the
    perturbation is a valid one (eps-ball, [0,1] box) and attacks the fp32 surrogate about as well as the strict path's —
    white-box loss increase >= 90 % of the strict one. It is off unless asked for."""
    net = _net("resnet50")
    x, y = _data(16)
    wrapped = tab.utils.wrap_model(net)
    ce = torch.nn.CrossEntropyLoss()

    def loss_of(d):
        with torch.no_grad():
            return float(ce(wrapped(x.cuda() + d), y.cuda()))
    strict = make_attack(tab, "mifgsm", net)
    assert strict.fast_mode == ""
    d0 = strict(x, y)
    res = {"clean": loss_of(torch.zeros_like(d0)), "strict": loss_of(d0)}
    for name, kw, mode in (("mifgsm", {}, "bnfold"), ("mifgsm", {}, "bf16"), ("vmifgsm", {"num_neighbor": 4, "epoch": 5}, "bnfold+bf16")):
        fast = make_attack(tab, name, net, **kw)
        fast.fast_mode = mode
        d1 = fast(x, y)
        assert d1.dtype == torch.float32 and float(d1.abs().max()) <= 16 / 255 + 1e-7
        adv = x.cuda() + d1
        assert float(adv.min()) >= 0.0 and float(adv.max()) <= 1.0
        res["fast_%s_%s" % (name, mode)] = loss_of(d1)
    REPORT["fast_mode"] = res
    gain_strict = res["strict"] - res["clean"]
    assert gain_strict > 0, res
    for mode in ("bnfold", "bf16"):
        assert res["fast_mifgsm_" + mode] - res["clean"] >= 0.9 * gain_strict, res
    assert res["fast_vmifgsm_bnfold+bf16"] > res["clean"], res


def test_dim_runs_inside_the_cuda_graph_with_the_reference_draws():
    """DIM / DI-TI-MI draw per call from the host generator. The graph path makes all `epoch` draws up front (same calls, same
    order), uploads per-iteration table records and lets the captured kernels read record *it (ta_dim_*_dyn): one captured graph,
    a new draw per replay, the same bits as the eager loop and as the reference restatement — also on a second batch."""
    net = _net()
    x, y = _data(2)
    for name, kw in (("dim", {"epoch": 4}), ("dim", {"epoch": 5, "diversity_prob": 1.0}), ("ditimi", {"epoch": 4}), ("siditimi", {"epoch": 2})):
        eager = make_attack(tab, name, net, **kw); eager.use_cuda_graph = False
        graph = make_attack(tab, name, net, **kw); graph.use_cuda_graph = True
        for seed, (xx, yy) in ((4, (x, y)), (9, _data(2, seed=7))):
            seed_all(seed); d_e = eager(xx, yy); r_e = torch.rand(1)
            seed_all(seed); d_g = graph(xx, yy); r_g = torch.rand(1)
            assert getattr(graph, "_graphs", None), getattr(graph, "_graph_error", None)
            assert torch.equal(d_e, d_g), (name, kw, seed)
            assert torch.equal(r_e, r_g)                  # the host generator was consumed identically
        REPORT["graph/" + name + "_" + "_".join("%s%s" % kv for kv in kw.items())] = {"bit_identical_to_eager": True}


def test_cuda_graph_is_refused_for_host_rng_transforms():
    """Admix draws a permutation per call and declares nothing: it must stay eager (and still equal an eager twin)"""
    net = _net()
    x, y = _data(2)
    atk = make_attack(tab, "admix", net, epoch=2)
    atk.use_cuda_graph = True
    seed_all(4); d = atk(x, y)
    assert not getattr(atk, "_graphs", None)
    ref = make_attack(tab, "admix", net, epoch=2)
    seed_all(4)
    assert torch.equal(d, ref(x, y))


def test_cli_attack_and_eval_modes(tmp_path):
    """main.py with the reference's flags on a tiny synthetic dataset (random-weight models: no network here): PNGs come out
    with the right shape and stay inside the epsilon ball around the inputs after uint8 truncation; --eval prints an ASR row."""
    import subprocess
    import sys
    from PIL import Image
    inp, out = tmp_path / "data", tmp_path / "adv"
    (inp / "images").mkdir(parents=True)
    rng = np.random.default_rng(0)
    names = ["img%d.png" % i for i in range(6)]
    for n in names:
        Image.fromarray(rng.integers(0, 256, (224, 224, 3), dtype=np.uint8)).save(inp / "images" / n)
    with open(inp / "labels.csv", "w") as f:
        f.write("filename,label,targeted_label\n")
        for i, n in enumerate(names):
            f.write("%s,%d,%d\n" % (n, i * 7, i * 7 + 1))
    cmd = [sys.executable, os.path.join(ROOT, "main.py"), "--input_dir", str(inp), "--output_dir", str(out), "--attack", "mifgsm",
           "--model", "resnet18", "--epoch", "2", "--batchsize", "4", "--random_weights", "--num_workers", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    for n in names:
        a = np.array(Image.open(out / n)).astype(np.int32)
        b = np.array(Image.open(inp / "images" / n)).astype(np.int32)
        assert a.shape == (224, 224, 3)
        assert np.abs(a - b).max() <= 17 and np.abs(a - b).max() >= 1      # eps = 16/255, truncation adds < 1
    r = subprocess.run(cmd + ["--eval"], capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "resnet50:" in r.stdout and r.stdout.strip().splitlines()[-1].startswith("|")
