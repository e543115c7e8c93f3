"""Host-side logic of transferattack_b200 on a box without a GPU: the Attack loop, the hook API and the plugin classes
are exercised with the kernels' stand-in (tests/oracle_backend.py, the C oracle behind the same backend interface) and
compared with (a) the eager-PyTorch restatement oracle/torch_ref.py and (b) end-to-end golden vectors produced by the
unmodified reference (tests/golden/e2e.npz).  The CUDA kernels themselves are checked in the -m gpu tests."""
import numpy as np
import pytest
import torch

import transferattack_b200 as tab
from transferattack_b200 import ops
from oracle import torch_ref
from oracle_backend import OracleBackend
from conftest import bits_equal, load_golden, n_diff_bits
from helpers import make_attack, seed_all, tiny_net


@pytest.fixture(autouse=True)
def oracle_backend():
    be = OracleBackend()
    ops._install_backend_for_tests(be)
    yield be
    ops._install_backend_for_tests(None)


@pytest.fixture(scope="module")
def E():
    return load_golden("e2e")


def _inputs(E):
    return torch.from_numpy(E["x"]), torch.from_numpy(E["y"])


def _host_matches_golden_host(E):
    """The e2e goldens depend on this host's CPU conv kernels; the stored first-forward logits are the fingerprint."""
    x, _ = _inputs(E)
    with torch.no_grad():
        l0 = torch_ref.ref_wrap_model(tiny_net(0))(x).numpy()
    return bits_equal(l0, E["logits0"])


MINE = {  # golden key -> (registry name, kwargs)
    "ifgsm": ("ifgsm", {}), "mifgsm": ("mifgsm", {}), "nifgsm": ("nifgsm", {}), "fgsm": ("fgsm", {}),
    "sim": ("sim", {}), "admix": ("admix", {}), "vmifgsm": ("vmifgsm", {"num_neighbor": 3}),
    "vnifgsm": ("vnifgsm", {"num_neighbor": 3}), "emifgsm": ("emifgsm", {}),
    "mifgsm_rs": ("mifgsm", {"random_start": True}), "mifgsm_targeted": ("mifgsm", {"targeted": True}),
}


def _run_mine(key, x, y, fuse=True, **extra):
    name, kw = MINE[key]
    atk = make_attack(tab, name, tiny_net(0), **kw, **extra)
    atk.fuse_update = fuse
    lab = torch.stack([y, (y + 1) % 10]) if kw.get("targeted") else y
    seed_all(2)
    return atk(x, lab), atk


def _run_ref(key, x, y):
    name, kw = MINE[key]
    kw = dict(kw)
    model = torch_ref.ref_wrap_model(tiny_net(0))
    atk = torch_ref.REF_ZOO[name](model, **kw)
    lab = torch.stack([y, (y + 1) % 10]) if kw.get("targeted") else y
    seed_all(2)
    return atk(x, lab)


@pytest.mark.parametrize("key", sorted(MINE))
def test_plugin_matches_torch_ref_bitwise(E, key):
    x, y = _inputs(E)
    d_ref = _run_ref(key, x, y)
    d_mine, atk = _run_mine(key, x, y)
    assert d_mine.shape == d_ref.shape and not d_mine.requires_grad
    assert bits_equal(d_mine.numpy(), d_ref.numpy()), n_diff_bits(d_mine.numpy(), d_ref.numpy())
    assert float(d_mine.abs().max()) <= atk.epsilon + 1e-8


@pytest.mark.parametrize("key", ["mifgsm", "ifgsm", "nifgsm", "sim"])
def test_fused_and_unfused_loops_agree(E, key, oracle_backend):
    x, y = _inputs(E)
    d_fused, _ = _run_mine(key, x, y, fuse=True)
    assert "fused_tail" in oracle_backend.calls
    oracle_backend.calls.clear()
    d_hooks, _ = _run_mine(key, x, y, fuse=False)
    assert "fused_tail" not in oracle_backend.calls and "momentum" in oracle_backend.calls
    assert bits_equal(d_fused.numpy(), d_hooks.numpy())


@pytest.mark.parametrize("mean_mode", ["torch", "exact"])
@pytest.mark.parametrize("name", ["mifgsm", "nifgsm", "tim"])
def test_normalize_fold_is_bit_identical(oracle_backend, name, mean_mode):
    """at the surrogate's native size the fused tail emits the normalised model input itself (and, in
    'exact' mean mode with the base get_grad, applies Normalize's adjoint too). Same ops in the same order → the same
    perturbation bit for bit as with the separate Normalize kernels and as the reference restatement."""
    gen = torch.Generator().manual_seed(11)
    x = torch.rand(2, 3, 224, 224, generator=gen); y = torch.randint(0, 10, (2,), generator=gen)
    epoch = 3
    atk = make_attack(tab, name, tiny_net(0), epoch=epoch)
    atk.mean_mode = mean_mode
    atk.fold_normalize = False
    oracle_backend.calls.clear()
    d_sep = atk(x, y)
    assert "fused_tail_nf" not in oracle_backend.calls
    n_sep = oracle_backend.calls.count("normalize")
    atk.fold_normalize = True
    oracle_backend.calls.clear()
    d_fold = atk(x, y)
    deferred = mean_mode == "exact" and name != "tim"          # TIM overrides get_grad: it needs the true gradient
    if name == "nifgsm":                                        # NI-FGSM overrides transform (look-ahead) → nothing to fold into
        assert "fused_tail_nf" not in oracle_backend.calls
    else:
        assert oracle_backend.calls.count("fused_tail_nf") == epoch and "fused_tail" not in oracle_backend.calls
        assert n_sep == 2 * epoch and oracle_backend.calls.count("normalize") == 1 + (0 if deferred else epoch)
    assert bits_equal(d_fold.numpy(), d_sep.numpy()), n_diff_bits(d_fold.numpy(), d_sep.numpy())
    if mean_mode == "torch":
        ref = torch_ref.REF_ZOO[name](torch_ref.ref_wrap_model(tiny_net(0)), epoch=epoch)(x, y)
        assert bits_equal(d_fold.numpy(), ref.numpy())


def test_gra_and_adaea_native_match_restatement(E, oracle_backend):
    """native GRA (ta_gra_update: decay indicator + tensor-step update in one launch) and AdaEA (ta_adaea_drf: the
    whole disparity-reduced filter in one launch) against the eager restatements of gradient/gra.py and ensemble/adaea.py
    (pinned to the live reference in tests/test_reference_live.py). GRA's ops are all bit-determined; AdaEA's filter enters
    through a 0/1 threshold on the map, so only pixels whose map value sits within rounding of the threshold could differ."""
    x, y = _inputs(E)
    kw = {"num_neighbor": 3, "epoch": 3}
    seed_all(2); ref = torch_ref.RefGRA(torch_ref.ref_wrap_model(tiny_net(0)), **kw)(x, y)
    oracle_backend.calls.clear()
    seed_all(2); d = make_attack(tab, "gra", tiny_net(0), **kw)(x, y)
    assert oracle_backend.calls.count("gra_update") == 3 and "update_linf" not in oracle_backend.calls
    assert bits_equal(d.numpy(), ref.numpy()), n_diff_bits(d.numpy(), ref.numpy())
    # the public hook alone (plugins that call get_decay_indicator themselves, e.g. the reference's fgsra.py)
    atk = make_attack(tab, "gra", tiny_net(0), **kw)
    M = torch.full_like(x, 1 / 0.94); cur = torch.randn_like(x); last = torch.randn_like(x)
    for lst in (0, last):
        lt = torch.zeros_like(cur) if isinstance(lst, int) else lst
        eq = (lt.sign() == cur.sign()).float()
        assert bits_equal(atk.get_decay_indicator(M, x, cur, lst, 0.94).numpy(), (M * (eq + (torch.ones_like(x) - eq) * 0.94)).numpy())
    nets = [tiny_net(0), tiny_net(3), tiny_net(5)]
    seed_all(4); ref = torch_ref.RefAdaEA(torch_ref.RefEnsemble([torch_ref.ref_wrap_model(n) for n in nets]), epoch=2)(x, y)
    oracle_backend.calls.clear()
    seed_all(4); d = make_attack(tab, "adaea", nets, epoch=2)(x, y)
    assert oracle_backend.calls.count("adaea_drf") == 2
    assert int((d != ref).sum()) <= 1e-5 * d.numel()


def test_pifgsm_native_matches_restatement(E, oracle_backend):
    """PI-FGSM on the kernels (ta_pi_cut_noise → ta_dwconv2d → ta_pi_update_linf) against the eager restatement
    of gradient/pifgsm.py. Every op is bit-exact except the 3x3 projection convolution, whose 8-term sums torch may add in
    another order; the result only enters through sign(), so a difference needs a sum that is exactly zero in one order and
    a rounding residue in the other — none on these inputs, and at most a handful per million is tolerated."""
    x, y = _inputs(E)
    for kw in ({}, {"decay": 1.0, "epoch": 4}, {"kern_size": 5, "epoch": 3}):
        ref = torch_ref.RefPIFGSM(torch_ref.ref_wrap_model(tiny_net(0)), **kw)(x, y)
        oracle_backend.calls.clear()
        atk = make_attack(tab, "pifgsm", tiny_net(0), **kw)
        d = atk(x, y)
        n = kw.get("epoch", 10)
        assert oracle_backend.calls.count("pi_cut_noise") == n and oracle_backend.calls.count("pi_update_linf") == n
        assert oracle_backend.calls.count("dwconv2d") == n
        bad = int((d != ref).sum())
        assert bad <= 1e-5 * d.numel(), (kw, bad)
        assert float(d.abs().max()) <= atk.epsilon + 1e-8


def test_normalize_fold_declines_what_it_cannot_fold(oracle_backend):
    x = torch.rand(2, 3, 32, 32); y = torch.tensor([1, 2])
    atk = make_attack(tab, "mifgsm", tiny_net(0), epoch=2)
    assert atk._fold_plan(x) is None                                         # Resize(224) is not a no-op at 32x32
    assert atk._fold_plan(torch.rand(1, 3, 224, 226)) is not None            # short side 224: Resize keeps the tensor
    assert atk._fold_plan(torch.rand(1, 3, 230, 226)) is None
    ens = make_attack(tab, "ens", [tiny_net(0), tiny_net(3)], epoch=2)
    assert ens._fold_plan(torch.rand(2, 3, 224, 224)) is None                # members normalise individually
    atk(x, y)
    assert "fused_tail_nf" not in oracle_backend.calls


@pytest.mark.parametrize("key", sorted(MINE) + ["ens"])
def test_against_reference_golden(E, key):
    if not _host_matches_golden_host(E):
        pytest.skip("this host's CPU conv kernels differ from the golden host's (fingerprint mismatch)")
    x, y = _inputs(E)
    if key == "ens":
        atk = make_attack(tab, "ens", [tiny_net(0), tiny_net(3)])
        seed_all(2)
        d = atk(x, y)
    else:
        d, _ = _run_mine(key, x, y)
    assert bits_equal(d.numpy(), E["delta_" + key]), n_diff_bits(d.numpy(), E["delta_" + key])


def test_torch_ref_against_reference_golden(E):
    if not _host_matches_golden_host(E):
        pytest.skip("fingerprint mismatch")
    x, y = _inputs(E)
    for key in sorted(MINE) + ["dim", "tim"]:
        name, kw = MINE.get(key, (key, {}))
        model = torch_ref.ref_wrap_model(tiny_net(0))
        atk = torch_ref.REF_ZOO[name](model, **kw)
        lab = torch.stack([y, (y + 1) % 10]) if kw.get("targeted") else y
        seed_all(2)
        assert bits_equal(atk(x, lab).numpy(), E["delta_" + key]), key


def test_l2_norm_path(E):
    x, y = _inputs(E)
    kw = dict(norm="l2", epsilon=1.0, alpha=0.2)
    atk = make_attack(tab, "mifgsm", tiny_net(0), **kw)
    seed_all(2)
    d = atk(x, y)
    ref = torch_ref.ref_mifgsm(torch_ref.ref_wrap_model(tiny_net(0)), **kw)
    seed_all(2)
    dr = ref(x, y)
    # norms: fp64 here vs torch's fp32 reductions → tolerance, not bits
    np.testing.assert_allclose(d.numpy(), dr.numpy(), rtol=0, atol=2e-6)
    assert float(d.view(d.shape[0], -1).norm(dim=1).max()) <= 1.0 + 1e-5


def test_trace_from_reference_replays_through_hooks(E):
    """Per-iteration (grad, momentum, delta) recorded inside the reference's own MI-FGSM run: feeding its inputs to this
    package's hooks must reproduce its outputs bit for bit (mean taken by the same torch op = strict mode)."""
    atk = make_attack(tab, "mifgsm", tiny_net(0))
    x, _ = _inputs(E)
    for i in range(int(E["trace_len"])):
        g = torch.from_numpy(E["trace%d_g" % i])
        m_in = torch.from_numpy(E["trace%d_m_in" % i]) if ("trace%d_m_in" % i) in E.files else 0
        m = atk.get_momentum(g, m_in)
        assert bits_equal(m.numpy(), E["trace%d_m_out" % i]), i
        d = atk.update_delta(torch.from_numpy(E["trace%d_d_in" % i]), x, m, atk.alpha)
        assert bits_equal(d.detach().numpy(), E["trace%d_d_out" % i]), i
        assert d.requires_grad and d.is_leaf


@pytest.mark.parametrize("name", ["dim", "tim", "ditimi", "siditimi"])
def test_dim_tim_single_step(E, name):
    """DIM's blend and TIM's conv are tolerance-level vs ATen (FMA contraction / summation order), so after sign() a few
    near-zero elements may flip: one iteration, <= 0.2 % of elements may differ, and only by 2*alpha."""
    x, y = _inputs(E)
    kw = dict(epoch=1)
    if name != "tim":
        kw["diversity_prob"] = 1.0
    atk = make_attack(tab, name, tiny_net(0), **kw)
    ref = torch_ref.REF_ZOO[name](torch_ref.ref_wrap_model(tiny_net(0)), **kw)
    seed_all(5); d = atk(x, y).numpy()
    seed_all(5); dr = ref(x, y).numpy()
    bad = np.abs(d - dr) > 1e-6
    assert bad.mean() <= 2e-3, bad.mean()
    assert np.abs(d - dr).max() <= 2 * atk.alpha + 1e-7


def test_dim_consumes_cpu_generator_like_reference():
    atk = make_attack(tab, "dim", tiny_net(0))
    ref = torch_ref.RefDIM(torch_ref.ref_wrap_model(tiny_net(0)))
    x = torch.rand(1, 3, 32, 32)
    for s in range(12):
        torch.manual_seed(s); p = atk.draw(32); a = torch.rand(1)
        torch.manual_seed(s); ref.transform(x); b = torch.rand(1)
        assert torch.equal(a, b)                       # same number of draws consumed
        assert (p is None) == (ref.last_params is None)
        if p is not None:
            assert tuple(p) == tuple(ref.last_params)


def test_hook_argument_tolerance():
    atk = make_attack(tab, "mifgsm", tiny_net(0))
    g = torch.randn(2, 3, 8, 8)
    for zero in (0, 0., 0.0):
        m = atk.get_momentum(g, zero, decay=0.3, foo="ignored")
        assert torch.equal(m, atk.get_momentum(g, 0))
    with pytest.raises(TypeError):
        atk.get_momentum(g, "zero")
    data = torch.rand(2, 3, 8, 8)
    delta = torch.zeros_like(data)
    a = atk.update_delta(delta, data, g, atk.alpha, projection=None)
    b = atk.update_delta(delta, data, g, torch.tensor(atk.alpha))
    c = atk.update_delta(delta, data, g, torch.full_like(data, atk.alpha))
    d = atk.update_delta(delta, data, g, torch.full((2, 1, 1, 1), atk.alpha))
    assert torch.equal(a, b) and torch.equal(a, c) and torch.equal(a, d)
    neg = atk.update_delta(delta, data, g, -atk.alpha)
    assert not torch.equal(a, neg)
    assert torch.equal(delta, torch.zeros_like(data))          # inputs untouched
    # non-contiguous views are accepted
    gt = g.permute(0, 1, 3, 2)
    # (torch's mean over a strided view may sum in another order: last-bit differences only)
    torch.testing.assert_close(atk.get_momentum(gt, 0), atk.get_momentum(gt.contiguous(), 0), rtol=1e-6, atol=0)


def test_errors_match_reference():
    with pytest.raises(Exception, match="Unsupported norm"):
        make_attack(tab, "mifgsm", tiny_net(0), norm="l1")
    with pytest.raises(Exception, match="Unsupported loss"):
        make_attack(tab, "mifgsm", tiny_net(0), loss="mse")
    with pytest.raises(Exception, match="Unspported attack algorithm"):
        tab.load_attack_class("nope")
    with pytest.raises(Exception, match="resize rate"):
        make_attack(tab, "dim", tiny_net(0), resize_rate=0.9)
    atk = make_attack(tab, "mifgsm", tiny_net(0), targeted=True)
    with pytest.raises(AssertionError):
        atk(torch.rand(3, 3, 8, 8), torch.zeros(3, dtype=torch.long))   # targeted needs [2, N] labels


def test_utils_star_export_surface():
    import transferattack_b200.utils as u
    ns = {}
    exec("from transferattack_b200.utils import *", ns)
    for name in ["torch", "nn", "models", "transforms", "Image", "np", "pd", "timm", "os", "img_height", "img_width", "img_max",
                 "img_min", "cnn_model_paper", "vit_model_paper", "cnn_model_pkg", "vit_model_pkg", "tgr_vit_model_list",
                 "generation_target_classes", "load_pretrained_model", "wrap_model", "save_images", "clamp",
                 "PreprocessingModel", "EnsembleModel", "AdvDataset"]:
        assert name in ns, name
    assert u.img_max == 1.0 and u.img_min == 0


def test_save_images_quantisation(tmp_path):
    from PIL import Image
    from transferattack_b200.utils import save_images
    M = load_golden("misc")
    save_images(str(tmp_path), torch.from_numpy(M["q_data"] + M["q_delta"]), ["a.png", "b.png"])
    u8 = np.stack([np.array(Image.open(tmp_path / f)) for f in ["a.png", "b.png"]])
    assert np.array_equal(u8, M["q_u8"])


def test_graph_capture_is_opt_in_per_hook_owner():
    """A CUDA graph replays what ran at capture time; a transform that flips a host coin per call must therefore never
    be captured. Every class defining a loop hook has to declare graph_safe itself — inheriting the flag is not enough."""
    ok = {n: make_attack(tab, n, [tiny_net(0), tiny_net(1)] if n in ("ens", "adaea") else tiny_net(0))._graph_ok() for n in tab.attack_zoo}
    assert ok == {"fgsm": True, "ifgsm": True, "mifgsm": True, "nifgsm": True, "tim": True, "sim": True, "ens": True,
                  "dim": True, "admix": False, "ditimi": True, "vmifgsm": False, "vnifgsm": False, "emifgsm": False, "pifgsm": False, "siditimi": True,
                  "gra": False, "adaea": False, "ssm": False}
    base = tab.load_attack_class("mifgsm")

    class Custom(base):                          # a user plugin overriding a hook without declaring anything
        def transform(self, x, **kw):
            return x if torch.rand(1) > 0.5 else x.flip(-1)
    assert not make_attack(tab, Custom, tiny_net(0))._graph_ok()

    class OnlyCtor(base):                        # overriding non-hook members keeps the parent's verdict
        def load_model(self, n):
            return super().load_model(n)
    assert make_attack(tab, OnlyCtor, tiny_net(0))._graph_ok()


def test_async_writer_and_prefetch_loader_match_the_serial_path(tmp_path):
    """SURVEY §8 f2 on CPU tensors: AsyncImageWriter produces the files save_images produces (same bytes in, same PNG out),
    errors surface at flush; PrefetchLoader yields the loader's batches unchanged, one ahead."""
    from PIL import Image
    from transferattack_b200.utils import AsyncImageWriter, PrefetchLoader, save_images
    g = torch.Generator().manual_seed(3)
    x = torch.rand(5, 3, 32, 32, generator=g); d = (torch.rand(5, 3, 32, 32, generator=g) - 0.5) * 0.1
    a, b = tmp_path / "sync", tmp_path / "async"
    a.mkdir(); b.mkdir()
    names = ["im%d.png" % i for i in range(5)]
    save_images(str(a), x, names, delta=d)
    with AsyncImageWriter(workers=3, max_pending=1) as w:
        w.submit(str(b), x[:2], names[:2], delta=d[:2])
        w.submit(str(b), x[2:], names[2:], delta=d[2:])
    for n in names:
        assert np.array_equal(np.array(Image.open(a / n)), np.array(Image.open(b / n)))
    w = AsyncImageWriter(workers=1)
    w.submit(str(tmp_path / "missing_dir"), x[:1], names[:1])
    with pytest.raises(Exception):
        w.flush()
    batches = [(x[i:i + 2], torch.arange(i, min(i + 2, 5)), names[i:i + 2]) for i in range(0, 5, 2)]
    got = list(PrefetchLoader(batches, "cpu"))
    assert len(got) == 3 and all(torch.equal(p[0], q[0]) and torch.equal(p[1], q[1]) and p[2] == q[2] for p, q in zip(got, batches))
    assert list(PrefetchLoader([], "cpu")) == []
