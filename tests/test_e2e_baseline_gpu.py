"""-m gpu: end-to-end parity AT BASELINE.json's OWN CONFIGURATIONS (VERDICT r1 "what's weak" §1), whole attacks through the
plugin API (kernels via the C-ABI) against oracle/torch_ref.py on the same GPU, same surrogate, seeds and cuDNN settings:

  configs[1]  MI-FGSM / ResNet-50 / B=64 / 10 iterations              bit-identical, 0 uint8 mismatches
  configs[2]  DI-TI-MI (DIM p=0.5 + TIM gaussian 15) / ResNet-50 / B=32 (the per-GPU share of 256/8) / 10 iterations
  configs[3]  VMI-FGSM N=20 / ViT-B/16 / 10 iterations                  (B=16: the full B=128 takes minutes per attack)
  configs[4]  ENS MI-FGSM {ResNet-50, ResNet-152, Inception-v3, ViT-B/16} on one device / 10 iterations

Tolerance (north_star): perturbation within 1e-5 abs fp32 and bit-identical after uint8 quantisation → asserted as
``n_gt_1e-5 == 0 and u8_mismatch == 0`` wherever the reference is itself deterministic; where the reference's own ops are
not run-to-run deterministic on CUDA (ATen's bilinear backward scatters with atomicAdd) the bound is the reference-vs-
reference floor measured in the same test. Reference lines: attack.py:67-102, dim.py:42-68, tim.py:68-73,
vmifgsm.py:42-97, utils.py:82-105, ens.py:31-36."""
import json
import os

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
    with open(os.path.join(ROOT, "gpurun_out", "e2e_parity_baseline.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


def _net(arch, seed=0):
    torch.manual_seed(seed)
    kw = {"aux_logits": True, "init_weights": False} if arch == "inception_v3" else {}
    return getattr(torchvision.models, arch)(weights=None, **kw).eval().cuda()


def _data(B, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, 224, 224, generator=g), torch.randint(0, 1000, (B,), generator=g)


def _stats(d, dr, x):
    d, dr = d.float().cpu(), dr.float().cpu()
    diff = (d - dr).abs()
    q = torch_ref.save_images_u8(x, d); qr = torch_ref.save_images_u8(x, dr)
    return {"max_abs": float(diff.max()), "n_gt_1e-5": int((diff > 1e-5).sum()), "numel": d.numel(),
            "u8_mismatch": int((q != qr).sum()), "bit_identical": bool(torch.equal(d, dr))}


def _run(fn, seed):
    seed_all(seed); torch.cuda.manual_seed_all(seed)
    out = fn()
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("mean_mode", ["torch", "aten"])
def test_config2_mifgsm_resnet50_b64_10iter_bit_identical(mean_mode):
    """BASELINE configs[1], the benchmarked configuration: every one of the 64 x 150,528 perturbation entries equal.
    'torch' = the in-kernel replay of ATen's reduction tree (one launch per tail); 'aten' = scale from ATen's own op."""
    net = _net("resnet50")
    x, y = _data(64)
    ref = torch_ref.ref_mifgsm(torch_ref.ref_wrap_model(net))
    dr = _run(lambda: ref(x, y), 2)
    atk = make_attack(tab, "mifgsm", net)
    atk.mean_mode = mean_mode
    d = _run(lambda: atk(x, y), 2)
    d2 = _run(lambda: atk(x, y), 2)               # second batch through the cached CUDA graph
    st = _stats(d, dr, x)
    st["cuda_graph"] = bool(getattr(atk, "_graphs", None))
    REPORT["config2/mifgsm_resnet50_b64_%s" % mean_mode] = st
    assert st["bit_identical"] and st["u8_mismatch"] == 0 and st["n_gt_1e-5"] == 0, st
    assert torch.equal(d, d2)
    assert st["cuda_graph"], getattr(atk, "_graph_error", None)


def test_config3_ditimi_resnet50_b32_10iter():
    """BASELINE configs[2] per-GPU share. DIM's forward reproduces ATen's kernels bit for bit and TIM's separable conv differs from
    cuDNN's at rounding level only; ATen's own bilinear backward is an atomicAdd scatter, so the bound is the reference's
    run-to-run floor (0 when it happens to be deterministic)."""
    net = _net("resnet50")
    x, y = _data(32)
    ref = torch_ref.REF_ZOO["ditimi"](torch_ref.ref_wrap_model(net), diversity_prob=0.5)
    dr = _run(lambda: ref(x, y), 6)
    dr2 = _run(lambda: ref(x, y), 6)
    atk = make_attack(tab, "ditimi", net, diversity_prob=0.5)
    d = _run(lambda: atk(x, y), 6)
    st, floor = _stats(d, dr, x), _stats(dr, dr2, x)
    REPORT["config3/ditimi_resnet50_b32"] = st
    REPORT["config3/ditimi_resnet50_b32_ref_vs_ref"] = floor
    assert float(d.abs().max()) <= 16 / 255 + 1e-7 and torch.isfinite(d).all()
    assert st["n_gt_1e-5"] <= floor["n_gt_1e-5"], (st, floor)
    assert st["u8_mismatch"] <= floor["u8_mismatch"], (st, floor)


@pytest.mark.parametrize("name", ["dim", "tim"])
def test_dim_tim_resnet50_b32_10iter(name):
    """The two halves of configs[2] on their own, 10 iterations, asserted (r1 only reported them)."""
    net = _net("resnet50")
    x, y = _data(32)
    kw = {} if name == "tim" else {"diversity_prob": 0.5}
    ref = torch_ref.REF_ZOO[name](torch_ref.ref_wrap_model(net), **kw)
    dr = _run(lambda: ref(x, y), 6)
    dr2 = _run(lambda: ref(x, y), 6)
    atk = make_attack(tab, name, net, **kw)
    d = _run(lambda: atk(x, y), 6)
    st, floor = _stats(d, dr, x), _stats(dr, dr2, x)
    REPORT["ten_iter_rn50_b32/" + name] = st
    REPORT["ten_iter_rn50_b32/" + name + "_ref_vs_ref"] = floor
    assert st["n_gt_1e-5"] <= floor["n_gt_1e-5"], (st, floor)
    assert st["u8_mismatch"] <= floor["u8_mismatch"], (st, floor)


def test_config4_vmifgsm_n20_vit_b16_10iter_bit_identical():
    """BASELINE configs[3]: N=20 neighbours, beta=1.5, ViT-B/16 (torchvision vit_b_16), 10 iterations. The neighbour noise is
    torch's own Philox stream (reproduced in the staging kernel), the accumulation order is the reference's."""
    net = _net("vit_b_16")
    x, y = _data(16)
    ref = torch_ref.REF_ZOO["vmifgsm"](torch_ref.ref_wrap_model(net), num_neighbor=20)
    dr = _run(lambda: ref(x, y), 3)
    atk = make_attack(tab, "vmifgsm", net, num_neighbor=20)
    d = _run(lambda: atk(x, y), 3)
    st = _stats(d, dr, x)
    REPORT["config4/vmifgsm_n20_vit_b16_b16"] = st
    assert st["bit_identical"] and st["u8_mismatch"] == 0, st


def test_config5_ens_four_members_one_device_10iter():
    """BASELINE configs[4] in the reference's own layout (all members on one device, utils.py:94-100). Inception-v3's wrapper
    resizes to 299 with antialiasing inside the autograd graph; its backward is torch's on both sides."""
    nets = [_net("resnet50", 0), _net("resnet152", 1), _net("inception_v3", 2), _net("vit_b_16", 3)]
    x, y = _data(16)
    ref = torch_ref.ref_mifgsm(torch_ref.RefEnsemble([torch_ref.ref_wrap_model(n) for n in nets]))
    dr = _run(lambda: ref(x, y), 4)
    dr2 = _run(lambda: ref(x, y), 4)
    atk = make_attack(tab, "ens", nets)
    d = _run(lambda: atk(x, y), 4)
    st, floor = _stats(d, dr, x), _stats(dr, dr2, x)
    REPORT["config5/ens4_one_device_b16"] = st
    REPORT["config5/ens4_one_device_b16_ref_vs_ref"] = floor
    assert st["n_gt_1e-5"] <= floor["n_gt_1e-5"] and st["u8_mismatch"] <= floor["u8_mismatch"], (st, floor)
    if floor["bit_identical"]:
        assert st["bit_identical"], st
