"""-m gpu: every entry point of libta_b200.so (called through the C-ABI, ctypes → raw device pointers) against
  (1) the golden vectors produced by the unmodified reference (tests/golden/*.npz),
  (2) the C oracle on seeded inputs (bit-exact wherever kernel and oracle share the op order),
  (3) size-independent properties at BASELINE sizes (B=64 x 3x224x224): adjoint identities, fused == unfused, bounds.
Tolerances are stated where used; everything else is bit-exact (NaN == NaN)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import bits_equal, load_golden, n_diff_bits, ulp_diff, ROOT

pytestmark = pytest.mark.gpu

EPS = 16 / 255
ALPHA = 1.6 / 255


@pytest.fixture(scope="module")
def be():
    from transferattack_b200 import ops
    ops._install_backend_for_tests(None)
    return ops.backend()


def cu(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


def test_library_and_device(be):
    from transferattack_b200 import _lib
    import ctypes
    lib = _lib.load()
    sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.ta_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi)) == 0
    assert (ma.value, mi.value) == (10, 0), "these kernels are built for sm_100a only"
    assert sm.value >= 100
    before = _lib.launch_count()
    be.add(torch.zeros(8, device="cuda"), torch.zeros(8, device="cuda"))
    assert _lib.launch_count() == before + 1


def test_cpu_tensor_is_rejected_loudly(be):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        be.add(torch.zeros(4), torch.zeros(4))


# ---------------------------------------------------------------------------------------------- golden: hooks
@pytest.fixture(scope="module")
def H():
    return load_golden("hooks")


@pytest.mark.parametrize("key,decay,first", [("mom_first", 1.0, True), ("mom_d1", 1.0, False), ("mom_d07", 0.7, False),
                                             ("mom_d0", 0.0, False)])
def test_momentum_golden(be, H, key, decay, first):
    out = be.momentum(cu(H["g"]), None if first else cu(H["m"]), cu(H["scale"]), decay)
    assert bits_equal(npy(out), H[key]), n_diff_bits(npy(out), H[key])


def test_momentum_nan_sample(be, H):
    scale = np.array(H["scale"]); scale[1] = 0.0
    out = npy(be.momentum(cu(H["gz"]), cu(H["m"]), cu(scale), 1.0))
    assert bits_equal(out, H["mom_nan"]) and np.isnan(out[1]).all()


def test_update_linf_golden(be, H):
    eps, alpha = float(H["eps"]), float(H["alpha"])
    d, x, m = cu(H["delta"]), cu(H["data"]), cu(H["mom_d1"])
    assert bits_equal(npy(be.update_linf(d, x, m, alpha, eps, 0, 1.0)), H["upd_linf"])
    assert bits_equal(npy(be.update_linf(d, x, m, -alpha, eps, 0, 1.0)), H["upd_linf_neg"])
    assert bits_equal(npy(be.update_linf(d, x, m, 0.0, eps, 0, 1.0, alpha_t=cu(H["alpha_t"]))), H["upd_linf_tensor"])
    assert bits_equal(npy(be.update_linf(d, x, cu(H["mom_nan"]), alpha, eps, 0, 1.0)), H["upd_nan"])


def test_fused_update_golden_strict(be, H):
    eps, alpha = float(H["eps"]), float(H["alpha"])
    g, m, d, x = cu(H["g"]), cu(H["m"]), cu(H["delta"]), cu(H["data"])
    m_out, d_out, xa = torch.empty_like(g), torch.empty_like(g), torch.empty_like(g)
    so = torch.empty(g.shape[0], device="cuda")
    be.fused_update_linf(g, m, m_out, d, d_out, x, xa, cu(H["scale"]), so, 1.0, alpha, eps, 0, 1.0)
    assert bits_equal(npy(m_out), H["mom_d1"]) and bits_equal(npy(d_out), H["upd_linf"])
    assert bits_equal(npy(xa), (H["data"] + H["upd_linf"]).astype(np.float32))
    assert bits_equal(npy(so), H["scale"])
    be.fused_update_linf(g, None, m_out, d, d_out, x, None, cu(H["scale"]), None, 1.0, alpha, eps, 0, 1.0)
    assert bits_equal(npy(m_out), H["mom_first"])


def test_l2_and_init_golden(be, H):
    eps = float(H["eps"])
    out = be.update_l2(cu(H["delta"] * np.float32(0.01)), cu(H["data"]), cu(H["g_l2"]), 0.01, eps, 0, 1.0)
    np.testing.assert_allclose(npy(out), H["upd_l2_small"], rtol=0, atol=2e-7)   # fp64 vs torch fp32 norm
    out = be.update_l2(cu(H["delta"]), cu(H["data"]), cu(H["g_l2"]), 2.0, eps, 0, 1.0)
    np.testing.assert_allclose(npy(out), H["upd_l2_big"], rtol=0, atol=2e-7)
    assert bits_equal(npy(out), oracle.update_l2(H["delta"], H["data"], H["g_l2"], 2.0, eps)) or \
        np.abs(npy(out) - oracle.update_l2(H["delta"], H["data"], H["g_l2"], 2.0, eps)).max() < 1e-7
    assert bits_equal(npy(be.clamp_box(cu(H["init_noise"]), cu(H["data"]), 0, 1.0)), H["init_linf"])
    out = be.init_l2_scale(cu(H["init_l2_normal"]), cu(H["init_l2_r"]), cu(H["data"]), eps, 0, 1.0)
    np.testing.assert_allclose(npy(out), H["init_l2"], rtol=0, atol=1e-8)


def test_stage_golden(be, H):
    assert bits_equal(npy(be.stage_add(cu(H["data"]), cu(H["delta"]))), H["x_adv"])
    assert bits_equal(npy(be.stage_add(cu(H["data"]), cu(H["delta"]), cu(H["m"]), float(H["ni_coef"]))), H["ni_x"])
    assert bits_equal(npy(be.stage_add(cu(H["x_adv"]), None, cu(H["m"]), float(H["ni_coef"]))), H["ni_x"])


def test_misc_golden(be):
    M = load_golden("misc")
    mean, std = cu(M["norm_mean"]), cu(M["norm_std"])
    assert bits_equal(npy(be.normalize(cu(M["norm_x"]), mean, std, True)), M["norm_y"])
    assert bits_equal(npy(be.normalize(cu(M["norm_gout"]), None, std, False)), M["norm_gin"])
    u8 = npy(be.quantize_u8(cu(M["q_data"]), cu(M["q_delta"]), True))
    assert np.array_equal(u8, M["q_u8"])
    u8c = npy(be.quantize_u8(cu(M["q_data"]), cu(M["q_delta"]), False))
    assert np.array_equal(u8c.transpose(0, 2, 3, 1), M["q_u8"])


def test_sim_admix_emi_golden(be):
    G = load_golden("sim_admix_emi")
    x = cu(G["sim_x"])
    S = int(G["sim_S"])
    assert bits_equal(npy(be.sim(x, S, True)), G["sim_y"])
    assert bits_equal(npy(be.sim(cu(G["sim_gout"]), S, False)), G["sim_gin"])
    S, A = int(G["admix_S"]), int(G["admix_A"])
    perm = torch.from_numpy(G["admix_perm"]).cuda()
    assert bits_equal(npy(be.admix(x, perm, float(G["admix_strength"]), S, A, True)), G["admix_y"])
    assert bits_equal(npy(be.admix(cu(G["admix_gout"]), None, 0.0, S, A, False)), G["admix_gin"])
    coef = [float(c) for c in G["emi_coef"]]
    assert bits_equal(npy(be.lin_sample(x, cu(G["emi_gbar"]), coef, True)), G["emi_y"])
    assert bits_equal(npy(be.lin_sample(x, None, coef, True)), G["emi_y0"])
    assert bits_equal(npy(be.lin_sample(cu(G["emi_gout"]), None, coef, False)), G["emi_gin"])


def test_vmi_golden(be):
    V = load_golden("vmi")
    N = int(V["N"])
    acc = None
    for k in range(N):
        xn = be.neighbor_stage(cu(V["data"]), cu(V["delta"]), cu(V["noises"][k]))
        assert bits_equal(npy(xn), V["x_near"][k])
        acc = be.accumulate(acc, cu(V["grads"][k]), first=(k == 0))
    var = be.variance_finalize(acc, cu(V["cur"]), N)
    assert bits_equal(npy(var), V["variance"])
    assert bits_equal(npy(be.add(cu(V["cur"]), var)), V["g_plus_v"])


# ---------------------------------------------------------------------------------------------- DIM
def _dim_cases():
    D = load_golden("dim")
    return D, sorted({k.rsplit("_", 1)[0] for k in D.files})


@pytest.fixture(params=[(2, 0, 0, 1), (4, 0, 0, 1), (4, 0, 1, 0), (3, 0, 0, 1), (3, 0, 1, 0), (1, 0, 0, 1), (1, 1, 1, 1), (0, 0, 0, 1)],
                ids=["default", "walk", "walk-wstab-rtpitch", "sep", "sep-wstab-rtpitch", "direct", "direct-gather-wstab", "fourpass"])
def dim_impl(request):
    """All generations of the DIM kernels must meet the same parity bar: the default (register-carried forward, separable-pass adjoint), the source-driven forward walk,
    the separable-pass kernels of csrc/dim_direct.cu in both directions (with compile-time and with run-time pitches), the
    second-generation kernels of the same file (forward with its tables as
    kernel parameters, adjoint = gather + scatter with the tables in the workspace; alternative: forward tables in the workspace,
    adjoint = independent gather) and the four-pass kernels of csrc/dim.cu."""
    from transferattack_b200 import _lib
    _lib.tune_set("dim.impl", request.param[0]); _lib.tune_set("dim.bwd", request.param[1]); _lib.tune_set("dim.fwdtab", request.param[2])
    _lib.tune_set("dim.sepconst", request.param[3])
    yield request.param
    _lib.tune_set("dim.impl", 2); _lib.tune_set("dim.bwd", 0); _lib.tune_set("dim.fwdtab", 0); _lib.tune_set("dim.sepconst", 1)


@pytest.mark.parametrize("tma", [1, 0])
def test_dim_forward(be, tma, dim_impl):
    from transferattack_b200 import _lib
    _lib.tune_set("dim.tma", tma)
    try:
        D, cases = _dim_cases()
        for c in cases:
            rnd, R, top, left, _ = [int(v) for v in D[c + "_params"]]
            for blend in (0, 1):
                _lib.tune_set("dim.blend", blend)
                out = npy(be.dim(cu(D[c + "_x"]), rnd, R, top, left, True))
                ref = oracle.dim_fwd(D[c + "_x"], rnd, R, top, left, blend=blend)
                assert bits_equal(out, ref), (c, blend, n_diff_bits(out, ref), np.abs(out - ref).max())   # same op order as the oracle
                np.testing.assert_allclose(out, D[c + "_y"], rtol=0, atol=3e-7, err_msg=c)               # ATen CPU golden: contraction level
    finally:
        _lib.tune_set("dim.tma", 1)
        _lib.tune_set("dim.blend", 1)


def test_dim_forward_bit_identical_to_torch_cuda(be, dim_impl):
    """The reference runs F.interpolate / F.pad / F.interpolate on the GPU; with the default blend (the FMA contraction of
    torch's own CUDA kernel) ta_dim_fwd reproduces that chain bit for bit. The adjoint is compared with autograd's
    (atomicAdd scatter) result at rounding level."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    for S, rnd, R, top, left in [(224, 235, 246, 5, 6), (224, 224, 246, 0, 21), (224, 245, 246, 1, 0), (64, 67, 70, 1, 2), (299, 310, 328, 7, 9), (30, 31, 33, 1, 1)]:
        x = torch.rand(3, 3, S, S, device="cuda", requires_grad=True)
        y = F.interpolate(F.pad(F.interpolate(x, size=[rnd, rnd], mode="bilinear", align_corners=False),
                                [left, R - rnd - left, top, R - rnd - top], value=0), size=[S, S], mode="bilinear", align_corners=False)
        out = be.dim(x, rnd, R, top, left, True)
        assert torch.equal(out, y.detach()), (S, rnd, int((out != y).sum()))
        g = torch.randn_like(y)
        (gin_ref,) = torch.autograd.grad(y, x, g)
        gin = be.dim(g, rnd, R, top, left, False)
        assert float((gin - gin_ref).abs().max()) <= 2e-6 * max(1.0, float(gin_ref.abs().max()))


def test_dim_backward(be, dim_impl):
    D, cases = _dim_cases()
    for c in cases:
        rnd, R, top, left, _ = [int(v) for v in D[c + "_params"]]
        gin = npy(be.dim(cu(D[c + "_gout"]), rnd, R, top, left, False))
        np.testing.assert_allclose(gin, oracle.dim_bwd(D[c + "_gout"], rnd, R, top, left), rtol=0, atol=2e-6, err_msg=c)
        np.testing.assert_allclose(gin, D[c + "_gin"], rtol=0, atol=3e-6, err_msg=c)


def test_dim_edge_geometries(be, dim_impl):
    rng = np.random.default_rng(3)
    for S, rate in [(224, 1.1), (299, 1.1), (64, 1.5), (33, 1.2), (16, 2.0)]:
        R = int(S * rate)
        x = rng.random((2, 3, S, S), dtype=np.float32)
        g = rng.standard_normal((2, 3, S, S)).astype(np.float32)
        for rnd, top, left in [(S, 0, 0), (R - 1, 0, 0), (R - 1, 1, 1), (S, R - S, R - S), ((S + R) // 2, 1, (R - (S + R) // 2))]:
            out = npy(be.dim(cu(x), rnd, R, top, left, True))
            ref = oracle.dim_fwd(x, rnd, R, top, left, blend=1)
            assert bits_equal(out, ref), (S, rnd, top, left, n_diff_bits(out, ref))
            gin = npy(be.dim(cu(g), rnd, R, top, left, False))
            np.testing.assert_allclose(gin, oracle.dim_bwd(g, rnd, R, top, left), rtol=0, atol=3e-6)


# ---------------------------------------------------------------------------------------------- TIM
def test_tim_conv(be):
    import transferattack_b200.input_transformation.tim as tim
    T = load_golden("tim")
    for key in sorted(k[:-7] for k in T.files if k.endswith("_kernel")):
        kt, ks = key.rstrip("0123456789"), int(key[len(key.rstrip("0123456789")):])
        k2d, kcol, krow = tim.make_kernel(kt, ks)
        assert bits_equal(k2d, T[key + "_kernel"])
        kc3, kr3 = np.stack([kcol] * 3), np.stack([krow] * 3)
        for tag in "abc":
            if key + "_" + tag + "_in" not in T.files:
                continue
            x = T[key + "_" + tag + "_in"]
            out2d = npy(be.dwconv2d(cu(x), cu(k2d.reshape(3, ks, ks))))
            assert bits_equal(out2d, oracle.dwconv2d(x, k2d)), (key, tag, "2d")
            outs = npy(be.dwconv2d_sep(cu(x), cu(kc3), cu(kr3)))
            assert bits_equal(outs, oracle.dwconv2d_sep(x, kc3, kr3)), (key, tag, "sep")
            # vs the reference's F.conv2d: summation order differs on both sides (N(0,1) inputs, weights sum to 1)
            np.testing.assert_allclose(out2d, T[key + "_" + tag + "_out"], rtol=0, atol=1e-6)
            np.testing.assert_allclose(outs, T[key + "_" + tag + "_out"], rtol=0, atol=1e-6)


def test_tim_generic_kernel_sizes(be):
    rng = np.random.default_rng(5)
    for ks in (1, 9, 11, 13, 21, 31):
        x = rng.standard_normal((1, 2, 40, 70)).astype(np.float32)
        k = rng.random((2, ks, ks), dtype=np.float32)
        assert bits_equal(npy(be.dwconv2d(cu(x), cu(k))), oracle.dwconv2d(x, k.reshape(2, 1, ks, ks))), ks
        kc, kr = rng.random((2, ks), dtype=np.float32), rng.random((2, ks), dtype=np.float32)
        assert bits_equal(npy(be.dwconv2d_sep(cu(x), cu(kc), cu(kr))), oracle.dwconv2d_sep(x, kc, kr)), ks


@pytest.mark.parametrize("ks", [3, 5, 7, 15])
def test_tim_sep_all_launch_paths_bit_identical(be, ks):
    """The separable convolution has several launch paths (the unrolled band walk with paired weights — the default for host factors and
    H % 32 == 0 —, the same walk fed from a warp-private cp.async ring, register-sliding fed from global memory or from bulk-TMA-staged
    shared memory, each with the factors as kernel parameters or loaded from device arrays, band height 32 / 56; two-pass
    band kernel; 32x32 tiles): all must equal the C oracle bit for bit,
    including ragged heights (last band partly / wholly outside the image) and channel-specific factors."""
    from transferattack_b200 import _lib
    rng = np.random.default_rng(ks)
    shapes = [(2, 3, 224, 224), (1, 2, 64, 32), (1, 3, 40, 36), (1, 1, 33, 32), (1, 2, 20, 32), (1, 1, 100, 512), (1, 2, 70, 40)]
    try:
        for shp in shapes:
            x = rng.standard_normal(shp).astype(np.float32)
            C = shp[1]
            k1c, k1r = rng.random(ks, dtype=np.float32), rng.random(ks, dtype=np.float32)
            shared = (np.stack([k1c] * C), np.stack([k1r] * C))
            distinct = (rng.random((C, ks), dtype=np.float32), rng.random((C, ks), dtype=np.float32))
            for kc, kr in (shared, distinct):
                want = oracle.dwconv2d_sep(x, kc, kr)
                for band, bh, f2 in ((5, 32, 1), (4, 32, 1), (3, 32, 1), (3, 56, 1), (3, 32, 0), (3, 56, 0), (2, 32, 0), (2, 56, 0), (1, 32, 0), (0, 32, 0)):
                    _lib.tune_set("tim.band", band); _lib.tune_set("tim.bh", bh); _lib.tune_set("tim.f2", f2)
                    got = npy(be.dwconv2d_sep(cu(x), cu(kc), cu(kr)))
                    assert bits_equal(got, want), (shp, "device factors", band, bh, f2)
                    for split, deep in (((0, 0), (1, 0), (0, 1)) if band == 4 else ((0, 0),)):
                        _lib.tune_set("tim.split", split); _lib.tune_set("tim.deep", deep)
                        got = npy(be.dwconv2d_sep(cu(x), cu(kc), cu(kr), host=(kc, kr)))
                        assert bits_equal(got, want), (shp, "host factors", band, bh, f2, split, deep)
                    _lib.tune_set("tim.split", 0); _lib.tune_set("tim.deep", 0)
    finally:
        _lib.tune_set("tim.band", 4); _lib.tune_set("tim.bh", 32); _lib.tune_set("tim.f2", 1)


def test_tim_sep_hw_refuses_what_it_cannot_serve(be):
    from transferattack_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(1, 3, 30, 30, device="cuda"); out = torch.empty_like(x)       # W % 4 != 0
    k = np.ones((3, 15), np.float32)
    rc = lib.ta_dwconv2d_sep_hw(x.data_ptr(), k.ctypes.data, k.ctypes.data, 15, out.data_ptr(), 1, 3, 30, 30, None)
    assert rc == _lib.TA_EUNSUPPORTED and "ta_dwconv2d_sep_hw" in _lib.last_error()


# ---------------------------------------------------------------------------------------------- reductions + fused
@pytest.mark.parametrize("B,shape", [(5, (3, 224, 224)), (3, (3, 20, 20)), (2, (37,)), (2, (3, 299, 299)), (1, (3, 512, 512)), (4, (8,))])
def test_abs_mean_exact(be, B, shape):
    rng = np.random.default_rng(B)
    g = (rng.standard_normal((B,) + shape) * 1e-3).astype(np.float32)
    got = npy(be.abs_mean(cu(g)))
    ref = oracle.abs_mean_per_sample(g)
    assert ulp_diff(got, ref).max() <= 1, (got, ref)


FUSED_SHAPES = [(5, (3, 224, 224)), (3, (3, 20, 20)), (2, (3, 299, 299)), (2, (37,)), (9, (3, 64, 64)), (1, (3, 512, 512))]
FUSED_TUNES = [dict(), {"fused.unroll": 1}, {"fused.unroll": 4}, {"fused.cluster": 4}, {"fused.cluster": 16, "fused.unroll": 1},
               {"fused.cluster": 1}, {"fused.cluster": 2, "fused.unroll": 4}]
FUSED_KEYS = {"fused.unroll": 2, "fused.cluster": 0}


def _torch_mean(g):
    """the reference's own op (attack.py:128) on the GPU: the bits TA_MEAN_TORCH must reproduce"""
    t = cu(g)
    return npy(t.abs().mean(dim=tuple(range(1, t.dim()))))


@pytest.mark.parametrize("B,shape", [(1, (3, 224, 224)), (2, (3, 224, 224)), (5, (3, 224, 224)), (64, (3, 224, 224)), (256, (3, 224, 224)),
                                     (16, (3, 64, 64)), (9, (3, 64, 64)), (31, (3, 224, 224)), (600, (3, 224, 224)), (8, (3, 384, 384)),
                                     (128, (1, 224, 224)), (64, (3, 300, 300))])
def test_abs_mean_torch_order_is_bit_identical_to_torch(be, B, shape):
    """TA_MEAN_TORCH replays the launch policy and summation tree of torch's CUDA mean kernel (csrc/aten_mean.cuh): the result
    must equal `g.abs().mean(dim=(1,2,3))` of the installed torch BIT FOR BIT, and the numpy restatement (oracle/aten_reduce.py)."""
    from transferattack_b200 import _lib
    from oracle import aten_reduce
    prop = torch.cuda.get_device_properties(0)
    for seed, scale in ((0, 1.0), (1, 1e-4), (2, 3e3)):
        rng = np.random.default_rng(B + seed)
        g = (rng.standard_normal((B,) + shape) * scale).astype(np.float32)
        got = be.abs_mean(cu(g), _lib.TA_MEAN_TORCH)
        assert got is not None, (B, shape)
        ref = _torch_mean(g)
        assert bits_equal(npy(got), ref), (B, shape, seed, ulp_diff(npy(got), ref).max())
        if g.size <= 40 * 150528:
            em = aten_reduce.emulate_numpy(np.abs(g).reshape(B, -1), prop.multi_processor_count, prop.max_threads_per_multi_processor)
            assert bits_equal(em, ref), (B, shape, seed)


@pytest.mark.parametrize("B,shape", [(1, (3, 224, 224)), (2, (3, 224, 224)), (5, (3, 224, 224)), (64, (3, 224, 224)), (65, (3, 224, 224)),
                                     (256, (3, 224, 224)), (16, (3, 64, 64)), (8, (3, 384, 384)), (128, (1, 224, 224)), (32, (3, 300, 300)),
                                     (600, (3, 224, 224)), (4, (4, 64, 64))])
def test_normalize_adjoint_with_column_sums_is_bit_identical_to_torch(be, B, shape):
    """ta_normalize_bwd_colsum = Normalize's adjoint (the bits of ta_normalize_bwd, i.e. of torchvision's div_ under autograd) that also
    leaves ATen's per-virtual-thread column values of |g|; ta_abs_mean_from_colsums finishes the mean from them: together they must
    equal `(gout / std).abs().mean(dim=(1,2,3))` of the installed torch BIT FOR BIT — the same contract as TA_MEAN_TORCH, with the
    gradient read once instead of twice."""
    C = shape[0]
    std = torch.tensor([0.229, 0.224, 0.225, 0.31][:C], device="cuda")
    n = int(np.prod(shape))
    S = be.colsum_size(B, n, torch.device("cuda", 0))
    assert S is not None, (B, shape)
    for seed, scale in ((0, 1.0), (1, 1e-4), (2, 3e3)):
        g = torch.randn((B,) + shape, device="cuda", generator=torch.Generator("cuda").manual_seed(B + seed)) * scale
        cs = torch.full((B * S,), float("nan"), device="cuda")
        out = torch.empty(B, device="cuda")
        gin = be.normalize_bwd_colsum(g, std, cs)
        assert gin is not None, (B, shape)
        ref = g / std.view(1, C, 1, 1)
        assert torch.equal(gin, ref) and torch.equal(gin, be.normalize(g, None, std, False)), (B, shape, seed)
        mu = be.abs_mean_from_colsums(cs, out, B, n)
        assert torch.equal(mu, ref.abs().mean(dim=(1, 2, 3))), (B, shape, seed)
        got = be.abs_mean(gin, __import__("transferattack_b200")._lib.TA_MEAN_TORCH)
        assert got is not None and torch.equal(mu, got)
        # the form the attack loop uses: the last CTA of every sample finishes the mean inside the launch; the ticket counters end at zero
        cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
        for _ in range(2):
            mu2 = torch.full((B,), float("nan"), device="cuda")
            gin2 = be.normalize_bwd_colsum(g, std, cs, mu2, cnt)
            assert gin2 is not None, (B, shape, S)
            assert torch.equal(gin2, ref) and torch.equal(mu2, mu) and int(cnt.abs().sum()) == 0, (B, shape, seed)


def test_normalize_adjoint_with_column_sums_declines_what_it_does_not_replay(be):
    from transferattack_b200 import ops
    assert be.colsum_size(2, 3 * 299 * 299, torch.device("cuda", 0)) is None           # n % 4 != 0
    assert be.colsum_size(4, 3 * 32 * 32, torch.device("cuda", 0)) is None             # one warp row per output: not restated
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda")
    assert ops.colsum_adjoint_ok(torch.zeros(2, 3, 299, 299, device="cuda"), std) is False
    assert ops.colsum_adjoint_ok(torch.zeros(6, 3, 224, 224, device="cuda"), std) is True


def test_abs_mean_torch_order_declines_what_it_does_not_replay(be):
    from transferattack_b200 import _lib
    # odd row length (head / tail elements take another ATen path), tiny rows (one warp row per output), more partials than fit
    for B, shape in [(2, (3, 299, 299)), (2, (37,)), (4, (8,)), (4, (3, 32, 32)), (3, (3, 20, 20)), (1, (3, 512, 512))]:
        assert be.abs_mean(torch.zeros((B,) + shape, device="cuda"), _lib.TA_MEAN_TORCH) is None, (B, shape)
    from transferattack_b200 import ops
    assert ops.aten_mean_replay_ok(torch.zeros(2, 3, 299, 299, device="cuda")) is False
    assert ops.aten_mean_replay_ok(torch.zeros(6, 3, 224, 224, device="cuda")) is True
    assert ops.aten_mean_replay_ok(torch.zeros(1, 3, 224, 224, device="cuda")) is True


@pytest.mark.parametrize("mean_mode", ["exact", "torch"])
@pytest.mark.parametrize("tune", FUSED_TUNES, ids=lambda t: ",".join("%s=%s" % kv for kv in t.items()) or "default")
def test_fused_update_in_kernel_mean(be, tune, mean_mode):
    """the cluster kernel for every tuning point, odd sizes (generic two-launch form), in place on momentum and delta.
    'exact': fp64 mean within 1 ulp of the correctly rounded one; 'torch': the mean is torch's own, bit for bit."""
    from transferattack_b200 import _lib
    mode = _lib.TA_MEAN_TORCH if mean_mode == "torch" else _lib.TA_MEAN_EXACT
    for k in FUSED_KEYS:
        _lib.tune_set(k, tune.get(k, FUSED_KEYS[k]))
    try:
        for B, shape in FUSED_SHAPES:
            rng = np.random.default_rng(B * 7 + len(shape))
            full = (B,) + shape
            g = (rng.standard_normal(full) * 1e-3).astype(np.float32)
            g.reshape(-1)[:7] = 0.0
            m = rng.standard_normal(full).astype(np.float32)
            x = rng.random(full, dtype=np.float32)
            d = ((rng.random(full, dtype=np.float32) * 2 - 1) * EPS).astype(np.float32)
            for has_m in (True, False):
                gm, dd = cu(m), cu(d)
                m_out, xa = torch.empty_like(gm), torch.empty_like(gm)
                so = torch.empty(B, device="cuda")
                # in place on momentum and delta, as the base loop uses it
                ok = be.fused_tail(cu(g), gm if has_m else None, gm if has_m else m_out, dd, dd, cu(x), xa, None, so, 0.9, ALPHA, EPS, 0, 1.0,
                                   mean_mode=mode)
                if not ok:                      # outside the replayed ATen launch family (or more columns per CTA than the forced
                    from oracle import aten_reduce        # cluster size holds): the caller passes torch's scale instead
                    prop = torch.cuda.get_device_properties(0)
                    n_el = int(np.prod(shape))
                    cfg = aten_reduce.config(B, n_el, prop.multi_processor_count, prop.max_threads_per_multi_processor)
                    cl = tune.get("fused.cluster", 0)
                    if cl <= 0:
                        cl = 1
                        while cl < 8 and n_el // (cl * 2) >= 2048:
                            cl *= 2
                    assert mean_mode == "torch" and (cfg is None or cfg["stride"] // cl > 3584), (B, shape, tune, cfg)
                    continue
                scale = npy(so)
                if mean_mode == "torch":
                    assert bits_equal(scale, _torch_mean(g)), (B, shape, tune)
                else:
                    assert ulp_diff(scale, oracle.abs_mean_per_sample(g)).max() <= 1
                mo, do, xo = oracle.fused_update_linf(g, m if has_m else None, d, x, scale, 0.9, ALPHA, EPS)
                got_m = npy(gm if has_m else m_out)
                assert bits_equal(got_m, mo), (B, shape, has_m, n_diff_bits(got_m, mo))
                assert bits_equal(npy(dd), do), (B, shape, has_m, n_diff_bits(npy(dd), do))
                assert bits_equal(npy(xa), xo), (B, shape, has_m)
    finally:
        for k in FUSED_KEYS:
            _lib.tune_set(k, FUSED_KEYS[k])


@pytest.mark.parametrize("mean_mode", ["scale", "exact", "torch"])
def test_fused_tail_addend_and_gbar(be, mean_mode):
    """ta_fused_tail's VMI / EMI options against the chain of reference ops (vmifgsm.py:87 `grad + variance`, emifgsm.py:97
    `grad / mean|grad|`): g' = g + v (one rounding), mean|g'|, momentum, update into a SECOND delta buffer (the old delta must
    survive for VMI's neighbours), next model input, and g'/mean as an extra output."""
    from transferattack_b200 import _lib
    for B, shape in [(5, (3, 224, 224)), (16, (3, 64, 64)), (3, (3, 20, 20))]:
        rng = np.random.default_rng(B)
        full = (B,) + shape
        g = (rng.standard_normal(full) * 1e-3).astype(np.float32)
        v = (rng.standard_normal(full) * 3e-4).astype(np.float32)
        m = rng.standard_normal(full).astype(np.float32)
        x = rng.random(full, dtype=np.float32)
        d = ((rng.random(full, dtype=np.float32) * 2 - 1) * EPS).astype(np.float32)
        for addend in (None, v):
            gsum = g if addend is None else oracle.add(g, addend)
            gm, dd = cu(m), cu(d)
            d_next, xa, gb = torch.empty_like(dd), torch.empty_like(dd), torch.empty_like(dd)
            so = torch.empty(B, device="cuda")
            if mean_mode == "scale":
                sc, mode = cu(_torch_mean(gsum)), _lib.TA_MEAN_EXACT
            else:
                sc, mode = None, (_lib.TA_MEAN_TORCH if mean_mode == "torch" else _lib.TA_MEAN_EXACT)
            ok = be.fused_tail(cu(g), gm, gm, dd, d_next, cu(x), xa, sc, so, 0.9, ALPHA, EPS, 0, 1.0, mean_mode=mode,
                               addend=cu(addend), gbar_out=gb)
            if mean_mode == "torch" and not ok:        # outside the replayed ATen launch family: the caller passes torch's scale
                assert be.abs_mean(cu(gsum), _lib.TA_MEAN_TORCH) is None, (B, shape)
                continue
            assert ok, (B, shape, _lib.last_error())
            scale = npy(so)
            if mean_mode == "exact":
                assert ulp_diff(scale, oracle.abs_mean_per_sample(gsum)).max() <= 1
            else:
                assert bits_equal(scale, _torch_mean(gsum)), (B, shape)
            mo, do, xo = oracle.fused_update_linf(gsum, m, d, x, scale, 0.9, ALPHA, EPS)
            tag = (B, shape, addend is not None)
            assert bits_equal(npy(gm), mo), tag
            assert bits_equal(npy(d_next), do) and bits_equal(npy(dd), d), tag          # old delta untouched
            assert bits_equal(npy(xa), xo), tag
            ref_gb = (gsum.reshape(B, -1) / scale.reshape(B, 1)).astype(np.float32).reshape(full)
            assert bits_equal(npy(gb), ref_gb), tag


@pytest.mark.parametrize("tune", [dict(), {"fused.unroll": 1}, {"fused.cluster": 4}], ids=["default", "unroll1", "cluster4"])
def test_fused_update_with_normalize_folded(be, tune):
    """ta_fused_update_linf_nf against the chain of reference ops it replaces (oracle.fused_update_linf_nf):
    strict (scale given) and exact (in-kernel mean) modes, gradient w.r.t. delta or w.r.t. the normalised input, first
    iteration (no momentum) and later ones, in place on momentum and delta."""
    from transferattack_b200 import _lib
    for k, v in {**FUSED_KEYS, **tune}.items():
        _lib.tune_set(k, v)
    try:
        for B, shape in [(5, (3, 224, 224)), (2, (3, 64, 64)), (3, (1, 32, 32)), (2, (4, 16, 16)), (2, (3, 226, 224))]:
            rng = np.random.default_rng(B + shape[0])
            full = (B,) + shape
            C = shape[0]
            mean = rng.random(C, dtype=np.float32); std = (0.2 + rng.random(C, dtype=np.float32)).astype(np.float32)
            g = (rng.standard_normal(full) * 1e-3).astype(np.float32)
            m = rng.standard_normal(full).astype(np.float32)
            x = rng.random(full, dtype=np.float32)
            d = ((rng.random(full, dtype=np.float32) * 2 - 1) * EPS).astype(np.float32)
            for wrt_xn in (False, True):
                for mmode in ("scale", "exact", "torch"):
                    strict = mmode == "scale"
                    for has_m in (True, False):
                        g_eff = oracle.normalize_bwd(g, std) if wrt_xn else g
                        gm, dd = cu(m), cu(d)
                        m_out, xn = torch.empty_like(gm), torch.empty_like(gm)
                        so = torch.empty(B, device="cuda")
                        sc = cu(oracle.abs_mean_per_sample(g_eff)) if strict else None
                        ok = be.fused_update_linf_nf(cu(g), gm if has_m else None, gm if has_m else m_out, dd, dd, cu(x), xn, sc, so,
                                                     0.9, ALPHA, EPS, 0, 1.0, mean, std, wrt_xn,
                                                     _lib.TA_MEAN_TORCH if mmode == "torch" else _lib.TA_MEAN_EXACT)
                        if mmode == "torch" and not ok:
                            assert be.abs_mean(cu(g_eff), _lib.TA_MEAN_TORCH) is None, (B, shape)
                            continue
                        assert ok
                        scale = npy(so)
                        if mmode == "torch":
                            assert bits_equal(scale, _torch_mean(g_eff)), (B, shape, wrt_xn)
                        else:
                            assert ulp_diff(scale, oracle.abs_mean_per_sample(g_eff)).max() <= (0 if strict else 1)
                        mo, do, xo, _ = oracle.fused_update_linf_nf(g, m if has_m else None, d, x, scale, 0.9, ALPHA, EPS, mean, std, wrt_xn)
                        tag = (B, shape, wrt_xn, strict, has_m)
                        assert bits_equal(npy(gm if has_m else m_out), mo), tag
                        assert bits_equal(npy(dd), do), tag
                        assert bits_equal(npy(xn), xo), tag
    finally:
        for k, v in FUSED_KEYS.items():
            _lib.tune_set(k, v)


def test_fused_update_nf_declines_unfoldable_shapes(be):
    for full in [(2, 3, 5, 5), (2, 5, 8, 8)]:          # plane % 4 != 0; more than 4 channels
        t = torch.zeros(full, device="cuda")
        so = torch.empty(full[0], device="cuda")
        C = full[1]
        assert be.fused_update_linf_nf(t, None, t.clone(), t.clone(), t.clone(), t, t.clone(), None, so, 1.0, ALPHA, EPS, 0, 1.0,
                                       [0.5] * C, [0.5] * C, False) is False


def test_fused_all_zero_gradient_sample(be):
    rng = np.random.default_rng(0)
    full = (3, 3, 32, 32)
    g = rng.standard_normal(full).astype(np.float32); g[1] = 0
    m = rng.standard_normal(full).astype(np.float32)
    x = rng.random(full, dtype=np.float32)
    d = np.zeros(full, np.float32)
    gm, dd = cu(m), cu(d)
    so = torch.empty(3, device="cuda")
    be.fused_update_linf(cu(g), gm, gm, dd, dd, cu(x), None, None, so, 1.0, ALPHA, EPS, 0, 1.0)
    assert npy(so)[1] == 0.0 and np.isnan(npy(gm)[1]).all()      # 0/0 → NaN momentum, as in the reference
    assert np.array_equal(npy(dd)[1], d[1])                      # sign(NaN) = 0 → delta does not move
    mo, do, _ = oracle.fused_update_linf(g, m, d, x, npy(so), 1.0, ALPHA, EPS)
    assert bits_equal(npy(gm), mo) and bits_equal(npy(dd), do)


def test_dim_dyn_kernels_equal_the_static_ones(be):
    """ta_dim_fwd_dyn / ta_dim_bwd_dyn (draw read from device memory at index *it, for CUDA-graph replay) against ta_dim_*_ws with
    the same draw: bit-identical in both directions, for every record incl. the identity coin, with the counter advanced on the
    device and clamped at the last record."""
    rng = np.random.default_rng(3)
    S, R = 224, 246
    x = cu(rng.random((2, 3, S, S), dtype=np.float32)); g = cu(rng.standard_normal((2, 3, S, S)).astype(np.float32))
    draws = [(235, 5, 6), None, (224, 0, 21), (245, 0, 0), (230, 16, 3)]
    host = be.dim_packs(draws, S, R)
    packs = host.cuda()
    it = torch.zeros(1, dtype=torch.int32, device="cuda")
    for i in range(len(draws) + 2):
        d = draws[min(i, len(draws) - 1)]
        f_dyn = be.dim_dyn(x, R, packs, len(draws), it, True)
        b_dyn = be.dim_dyn(g, R, packs, len(draws), it, False)
        if d is None:
            assert torch.equal(f_dyn, x) and torch.equal(b_dyn, g)
        else:
            assert torch.equal(f_dyn, be.dim(x, d[0], R, d[1], d[2], True)), (i, d)
            assert torch.equal(b_dyn, be.dim(g, d[0], R, d[1], d[2], False)), (i, d)
        be.counter_add(it, delta=1)
    assert int(it.item()) == len(draws) + 2
    be.counter_add(it, set_to=0)
    assert int(it.item()) == 0


# ---------------------------------------------------------------------------------------------- GRA / AdaEA
@pytest.mark.parametrize("shape", [(4, 3, 224, 224), (2, 3, 17, 19), (64, 3, 224, 224)])
def test_gra_update_matches_oracle(be, shape):
    """ta_gra_update (gra.py:74-93 + 149) — every op has a determined order: bit-exact, incl. NaN / zero momentum entries, the
    first iteration's python-0 `last`, and in place on M and delta"""
    rng = np.random.default_rng(sum(shape))
    M = (1 / 0.94 * 0.94 ** rng.integers(0, 4, shape)).astype(np.float32)
    cur = rng.standard_normal(shape).astype(np.float32); last = rng.standard_normal(shape).astype(np.float32)
    cur.reshape(-1)[:5] = 0.0; last.reshape(-1)[3:8] = 0.0; cur.reshape(-1)[9] = np.nan; last.reshape(-1)[10] = np.nan
    x = rng.random(shape, dtype=np.float32)
    d = ((rng.random(shape, dtype=np.float32) * 2 - 1) * EPS).astype(np.float32)
    for lst in (None, last):
        m1, d1 = be.gra_update(cu(M), cu(lst), cu(cur), 0.94, ALPHA, cu(d), cu(x), EPS, 0.0, 1.0)
        om, od = oracle.gra_update(M, lst, cur, 0.94, ALPHA, d, x, EPS)
        assert bits_equal(npy(m1), om) and bits_equal(npy(d1), od), (shape, lst is None)
        if lst is not None:      # against the reference's own ops on the GPU
            tM, tl, tc, td, tx = cu(M), cu(last), cu(cur), cu(d), cu(x)
            eq = (tl.sign() == tc.sign()).float()
            M2 = tM * (eq + (torch.ones_like(td) - eq) * 0.94)
            d2 = torch.clamp(td + (M2 * ALPHA) * tc.sign(), -EPS, EPS)
            d2 = torch.min(torch.max(d2, 0 - tx), 1.0 - tx)
            assert bits_equal(npy(m1), npy(M2)) and bits_equal(npy(d1), npy(d2))


@pytest.mark.parametrize("K,shape", [(4, (3, 3, 224, 224)), (2, (2, 3, 17, 19)), (3, (8, 3, 64, 64)), (8, (2, 3, 32, 32))])
def test_adaea_drf_matches_oracle_and_torch(be, K, shape):
    """ta_adaea_drf (adaea.py:115-136, 74-76, 82): the map within 2e-6 of the C oracle and of torch's own op chain (torch's order
    inside its 3-element norms / dot products is not specified), the thresholded product equal wherever the map is not within
    2e-6 of the threshold."""
    import torch.nn.functional as F
    rng = np.random.default_rng(K)
    grads = [(rng.standard_normal(shape) * 10.0 ** rng.integers(-6, 0)).astype(np.float32) for _ in range(K)]
    grads[0][0, :, 0, :5] = 0.0                                   # all-zero pixels: normalize → 0, cosine → 0
    grad = rng.standard_normal(shape).astype(np.float32)
    thr = -0.3
    out, mp = be.adaea_drf([cu(g) for g in grads], thr, cu(grad), want_map=True)
    omp, oout = oracle.adaea_drf(grads, thr, grad)
    assert np.abs(npy(mp) - omp).max() <= 2e-6
    tg = [cu(g) for g in grads]
    B, _, H, W = shape
    pair = torch.zeros(K, K, B, H, W, device="cuda"); rows = torch.zeros(K, B, H, W, device="cuda")
    cos = torch.nn.CosineSimilarity(dim=1, eps=1e-8)
    for i in range(K):
        for j in range(i + 1, K):
            pair[i][j] = cos(F.normalize(tg[i], dim=1), F.normalize(tg[j], dim=1))
        if i < K - 1:
            rows[i] = (pair[i, :].sum(dim=0) + pair[:, i].sum(dim=0)) / (K - 1)
    tmap = rows.mean(dim=0).view(B, 1, H, W)
    assert float((mp - tmap).abs().max()) <= 2e-6
    mask = (tmap >= thr).float()
    safe = ((tmap - thr).abs() > 2e-6).expand(-1, shape[1], -1, -1)
    assert torch.equal(out[safe], (cu(grad) * mask)[safe])
    assert bits_equal(npy(out)[np.abs(omp - thr).repeat(shape[1], 1) > 2e-6], oout[np.abs(omp - thr).repeat(shape[1], 1) > 2e-6])


# ---------------------------------------------------------------------------------------------- PI-FGSM
@pytest.mark.parametrize("shape", [(4, 3, 224, 224), (2, 3, 17, 19), (1, 3, 8, 8)])
def test_pifgsm_kernels_match_oracle(be, shape):
    rng = np.random.default_rng(sum(shape))
    coef, gamma = 10.0 * ALPHA, 16.0 / 255
    m = rng.standard_normal(shape).astype(np.float32); m.reshape(-1)[:5] = 0.0; m.reshape(-1)[5] = np.nan
    amp = (rng.standard_normal(shape) * 0.2).astype(np.float32)
    x = rng.random(shape, dtype=np.float32)
    d = ((rng.random(shape, dtype=np.float32) * 2 - 1) * EPS).astype(np.float32)
    for a0 in (None, amp):
        a1, cut = be.pi_cut_noise(cu(a0), cu(m), coef, EPS)
        oa, oc = oracle.pi_cut_noise(a0, m, coef, EPS)
        assert bits_equal(npy(a1), oa) and bits_equal(npy(cut), oc), (shape, a0 is None)
        k = np.ones((3, 3, 3), np.float32) / 8; k[:, 1, 1] = 0
        conv = be.dwconv2d(cut, cu(k))
        assert bits_equal(npy(conv), oracle.dwconv2d(oc, k.reshape(3, 1, 3, 3)))
        a2, dn = be.pi_update_linf(cu(d), cu(x), cu(m), conv, a1, coef, gamma, EPS, 0, 1.0)
        oa2, od = oracle.pi_update_linf(d, x, m, npy(conv), oa, coef, gamma, EPS)
        assert bits_equal(npy(a2), oa2) and bits_equal(npy(dn), od), (shape, a0 is None)
        assert np.nanmax(np.abs(npy(dn))) <= EPS + 1e-8


# ---------------------------------------------------------------------------------------------- VMI noise in the kernel
@pytest.mark.parametrize("shape", [(64, 3, 224, 224), (4, 3, 224, 224), (2, 3, 299, 299), (1, 3, 17, 19), (1000,), (3, 5, 7)])
def test_neighbor_stage_philox_reproduces_torch_uniform(be, shape):
    """ta_neighbor_stage_philox must put exactly the numbers of `zeros_like(delta).uniform_(-r, r)` (vmifgsm.py:50) at exactly
    the same elements, leave torch's device generator where that call would have left it, and equal the numpy restatement."""
    from oracle import philox as P
    r = 1.5 * EPS
    data = torch.rand(shape, device="cuda"); delta = (torch.rand(shape, device="cuda") * 2 - 1) * EPS
    look = torch.randn(shape, device="cuda")
    gen = torch.cuda.default_generators[0]                               # (populated once CUDA is initialised)
    for seed in (0, 1234567, 2 ** 40 + 17):
        for warm in (0, 3):
            torch.cuda.manual_seed(seed)
            for _ in range(warm):
                torch.rand(1000, device="cuda")                          # move the offset off zero
            state = gen.get_state()
            off0 = gen.get_offset()
            want_noise = torch.zeros_like(delta).uniform_(-r, r)
            off_torch = gen.get_offset()
            follow_torch = torch.rand(5, device="cuda")
            gen.set_state(state)
            noise = torch.empty_like(delta)
            out = be.neighbor_stage_philox(data, delta, -r, r, noise_out=noise)
            assert gen.get_offset() == off_torch                         # generator advanced identically
            assert torch.equal(torch.rand(5, device="cuda"), follow_torch)
            assert torch.equal(noise, want_noise), (shape, seed, warm, int((noise != want_noise).sum()))
            assert torch.equal(out, (data + delta) + want_noise)
            T, inc = be.torch_uniform_policy(delta.numel())
            assert inc == off_torch - off0
            if delta.numel() <= 1_000_000:
                ref = P.torch_uniform(delta.numel(), seed, off0, -r, r, T).reshape(shape)
                assert bits_equal(npy(noise), ref), (shape, seed)
    # with the Nesterov look-ahead term of VNI-FGSM
    gen.set_state(state)
    out = be.neighbor_stage_philox(data, delta, -r, r, look=look, coef=0.01)
    assert torch.equal(out, ((data + delta) + want_noise) + 0.01 * look)


# ---------------------------------------------------------------------------------------------- BASELINE-size properties
def test_full_size_fused_equals_unfused_and_bounds(be):
    torch.manual_seed(0)
    B = 64
    g = torch.randn(B, 3, 224, 224, device="cuda") * 1e-4
    m = torch.randn_like(g)
    x = torch.rand_like(g)
    d = (torch.rand_like(g) * 2 - 1) * EPS
    m2, d2, x2, so = torch.empty_like(g), torch.empty_like(g), torch.empty_like(g), torch.empty(B, device="cuda")
    be.fused_update_linf(g, m, m2, d, d2, x, x2, None, so, 1.0, ALPHA, EPS, 0, 1.0)
    scale = be.abs_mean(g)
    assert ulp_diff(npy(so), npy(scale)).max() <= 1          # two fp64 reduction trees, same value up to a final-rounding tie
    scale = so
    m1 = be.momentum(g, m, scale, 1.0)
    d1 = be.update_linf(d, x, m1, ALPHA, EPS, 0, 1.0)
    x1 = be.stage_add(x, d1)
    assert torch.equal(m1, m2) and torch.equal(d1, d2) and torch.equal(x1, x2)
    m3, d3, x3 = torch.empty_like(g), torch.empty_like(g), torch.empty_like(g)
    be.fused_update_linf(g, m, m3, d, d3, x, x3, scale, None, 1.0, ALPHA, EPS, 0, 1.0)   # strict path
    assert torch.equal(m1, m3) and torch.equal(d1, d3) and torch.equal(x1, x3)
    assert float(d2.abs().max()) <= np.float32(EPS)
    assert float(x2.min()) >= 0.0 and float(x2.max()) <= 1.0 + 1e-7
    # projection is idempotent
    assert torch.equal(be.clamp_box(d2, x, 0, 1.0), d2)
    # a sample-wise check of the mean against the oracle (order-independent up to the last bit)
    ref = oracle.abs_mean_per_sample(npy(g[:4]))
    assert ulp_diff(npy(so[:4]), ref).max() <= 1


def test_full_size_adjoint_identities(be):
    """<A x, g> == <x, A^T g> for the staging kernels at B=64 (fp64 dot products, relative 1e-5)."""
    torch.manual_seed(1)
    B = 64
    x = torch.rand(B, 3, 224, 224, device="cuda")

    def dot(a, b):
        return float((a.double() * b.double()).sum())

    y = be.sim(x, 5, True); g = torch.randn_like(y)
    assert abs(dot(y, g) - dot(x, be.sim(g, 5, False))) <= 1e-5 * abs(dot(y, g)) + 1e-3
    y = be.dim(x, 235, 246, 5, 7, True); g = torch.randn_like(y)
    assert abs(dot(y, g) - dot(x, be.dim(g, 235, 246, 5, 7, False))) <= 1e-5 * abs(dot(y, g)) + 1e-2
    coef = [float(np.float32(c * ALPHA)) for c in np.linspace(-7, 7, 11)]
    xs = x[:16]
    y = be.lin_sample(xs, None, coef, True); g = torch.randn_like(y)
    assert abs(dot(y, g) - dot(xs, be.lin_sample(g, None, coef, False))) <= 1e-5 * abs(dot(y, g)) + 1e-3


def test_full_size_tim_properties(be):
    import transferattack_b200.input_transformation.tim as tim
    k2d, kcol, krow = tim.make_kernel("gaussian", 15)
    kc3, kr3 = cu(np.stack([kcol] * 3)), cu(np.stack([krow] * 3))
    g = torch.randn(64, 3, 224, 224, device="cuda")
    a = be.dwconv2d_sep(g, kc3, kr3)
    b = be.dwconv2d(g, cu(k2d.reshape(3, 15, 15)))
    assert float((a - b).abs().max()) <= 2e-6            # separable vs direct: fp32 re-association only
    ones = torch.ones(2, 3, 224, 224, device="cuda")
    o = be.dwconv2d_sep(ones, kc3, kr3)
    assert float((o[:, :, 7:-7, 7:-7] - 1).abs().max()) <= 1e-6      # weights sum to 1 away from the zero padding
    # linearity
    g2 = torch.randn_like(g)
    lhs = be.dwconv2d_sep(be.add(g, g2), kc3, kr3)
    rhs = be.add(a, be.dwconv2d_sep(g2, kc3, kr3))
    assert float((lhs - rhs).abs().max()) <= 1e-5
    # one sample against the oracle, bit-exact
    assert bits_equal(npy(a[:1]), oracle.dwconv2d_sep(npy(g[:1]), npy(kc3), npy(kr3)))
