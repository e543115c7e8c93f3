"""CPU: the index logic of the second-generation TIM / DIM kernels (tests/kernel_models.py: band geometry, descriptor tables,
accumulator rotation, retire logic — mirrored step by step from csrc/dwconv.cu and csrc/dim_direct.cu) against the C oracle.
The CUDA kernels themselves are compared with the same oracle in the -m gpu tests."""
import numpy as np
import pytest

import oracle
import kernel_models as KM
from conftest import bits_equal


@pytest.mark.parametrize("ks,H,W,bhr", [(15, 64, 32, 32), (3, 40, 36, 32), (7, 33, 32, 32), (5, 70, 40, 56), (15, 20, 32, 32)])
def test_register_sliding_convolution_model(ks, H, W, bhr):
    """rotating k x 4 accumulator file, warm-up rows, ragged last band, zero padding by predicate: every output row stored
    exactly once and bit-identical to orc_dwconv2d_sep"""
    rng = np.random.default_rng(ks * 1000 + H)
    g = rng.standard_normal((1, 2, H, W)).astype(np.float32)
    kc = rng.random((2, ks)).astype(np.float32); kr = rng.random((2, ks)).astype(np.float32)
    assert bits_equal(KM.rs_conv_model(g, kc, kr, bhr), oracle.dwconv2d_sep(g, kc, kr))


@pytest.mark.parametrize("ks,H,W,bhr", [(15, 64, 32, 32), (15, 96, 64, 32), (7, 64, 32, 32), (3, 64, 36, 32), (5, 32, 40, 32)])
def test_unrolled_band_walk_model(ks, H, W, bhr):
    """dwconv_sep_rg2_kernel / dwconv_sep_rg3_kernel (the TIM kernel that ships in round 2): paired row weights with scalar end taps,
    the tap-exact column pass (a halo row feeds only the output rows that exist; tap 0 starts from +0; no slot is read before it
    was started), skipped out-of-image rows in the first / last band, zero reads outside the row — every output row stored
    exactly once and bit-identical to orc_dwconv2d_sep (a single band, H == bhr, has both halos outside the image)"""
    rng = np.random.default_rng(ks * 77 + H)
    g = rng.standard_normal((1, 2, H, W)).astype(np.float32)
    k1c = rng.random(ks).astype(np.float32); k1r = rng.random(ks).astype(np.float32)
    kc = np.stack([k1c] * 2); kr = np.stack([k1r] * 2)
    assert bits_equal(KM.rg2_conv_model(g, kc, kr, bhr), oracle.dwconv2d_sep(g, kc, kr))


@pytest.mark.parametrize("B,shape", [(64, (3, 32, 32 * 4)), (5, (3, 64, 64)), (2, (3, 64, 64)), (16, (4, 32, 64))])
def test_adjoint_with_block_trees_model_equals_the_aten_restatement(B, shape):
    """normalize_bwd_colsum_kernel<FINISH> step by step in numpy: the adjoint g / std[channel] per 128-bit vector in ATen's
    thread <-> data mapping, block_x_reduce / block_y_reduce per CTA, the final tree over the per-block partials — must equal
    oracle/aten_reduce.py's restatement of torch's `(g / std).abs().mean(dim=(1,2,3))` bit for bit (that restatement is pinned against
    torch on the GPU box) and the plain quotient for the gradient."""
    from oracle import aten_reduce
    n = int(np.prod(shape))
    cfg = aten_reduce.config(B, n)
    assert cfg is not None and cfg["bw"] * cfg["bh"] == 512
    rng = np.random.default_rng(B * 31 + n)
    g = (rng.standard_normal((B,) + shape) * 10.0 ** rng.integers(-3, 2)).astype(np.float32)
    std = np.asarray([0.229, 0.224, 0.225, 0.31][:shape[0]], np.float32)
    gin, mean = KM.colsum_adjoint_model(g, std, cfg)
    want = (g / std.reshape(1, -1, 1, 1)).astype(np.float32)
    assert bits_equal(gin, want)
    assert bits_equal(mean, aten_reduce.emulate_numpy(np.abs(want).reshape(B, n)))


DIM_GEOMETRIES = [(32, 33, 35, 1, 2), (32, 32, 35, 0, 3), (32, 34, 35, 1, 0), (20, 21, 22, 0, 0), (40, 43, 44, 1, 1),
                  (24, 24, 48, 10, 20), (33, 36, 36, 0, 0), (48, 52, 52, 0, 0), (16, 30, 40, 5, 5)]


@pytest.mark.parametrize("S,rnd,R,top,left", DIM_GEOMETRIES)
def test_dim_direct_models(S, rnd, R, top, left):
    """host tap / inverse tables and band table; forward: zero row + zero column as the padding → bit-identical to orc_dim_fwd
    (blend 1); the three adjoint forms (separable passes = the default, scatter, gather): every destination written exactly once, no
    read outside what a previous pass wrote, equal to the fp64 scatter oracle to rounding"""
    rng = np.random.default_rng(S * 100 + rnd)
    x = rng.random((1, S, S)).astype(np.float32); g = rng.standard_normal((1, S, S)).astype(np.float32)
    assert bits_equal(KM.dim_fwd_model(x, rnd, R, top, left), oracle.dim_fwd(x[None], rnd, R, top, left, blend=1)[0])
    want = oracle.dim_bwd(g[None], rnd, R, top, left)[0]
    for model in (KM.dim_bwd_sep_model, KM.dim_bwd_scatter_model, KM.dim_bwd_gather_model):
        got = model(g, rnd, R, top, left)
        assert not np.isnan(got).any()
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)


@pytest.mark.parametrize("B,n,cl", [(64, 150528, 8), (5, 150528, 8), (2, 12288, 1), (9, 12288, 4), (256, 150528, 8), (16, 12288, 16),
                                    (1, 150528, 8), (600, 150528, 8)])
def test_aten_mean_kernel_model_equals_the_aten_restatement(B, n, cl):
    """The thread/lane mapping of csrc/aten_mean.cuh (vector columns per CTA, x tree per block row in one warp, y tree per virtual
    block in one thread, final tree) executed step by step in numpy reproduces oracle/aten_reduce.py's restatement of ATen's
    reduction bit for bit; the
    restatement itself is pinned against torch on the GPU box (tests/test_kernels_gpu.py, tools/diag_aten_mean.py). Also: the C
    policy (ta_aten_mean_policy, host-only) equals the Python restatement of setReduceConfig over a grid of shapes."""
    import ctypes
    from oracle import aten_reduce
    from kernel_models import aten_mean_kernel_model
    cfg = aten_reduce.config(B, n)
    rng = np.random.default_rng(B + n)
    x = np.abs(rng.standard_normal(n) * 10.0 ** rng.integers(-4, 3)).astype(np.float32)
    tot = aten_mean_kernel_model(x, cfg, cl)
    mine = np.float32(tot * np.float32(np.float32(B) / np.float32(B * n)))
    ref = aten_reduce.emulate_numpy(np.repeat(x[None], B, 0))
    assert mine == ref[0] and (ref == ref[0]).all()


def test_aten_mean_policy_c_equals_python():
    import ctypes
    from oracle import aten_reduce
    from transferattack_b200 import _lib
    lib = _lib.load()
    for sm, mt in ((148, 2048), (132, 2048), (108, 2048), (84, 1536)):
        for B in (1, 2, 3, 4, 7, 8, 15, 16, 31, 32, 64, 65, 128, 256, 512, 592, 593, 600, 1024):
            for n in (8, 31, 32, 37, 128, 512, 1024, 1200, 2048, 3072, 12288, 50176, 150528, 268203, 268204, 442368, 786432):
                bw, bh, cpo = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
                rc = lib.ta_aten_mean_policy(B, n, sm, mt, ctypes.byref(bw), ctypes.byref(bh), ctypes.byref(cpo))
                cfg = aten_reduce.config(B, n, sm, mt)
                if cfg is None:
                    assert rc == _lib.TA_EUNSUPPORTED, (B, n, sm)
                else:
                    assert rc == 0 and (bw.value, bh.value, cpo.value) == (cfg["bw"], cfg["bh"], cfg["cpo"]), (B, n, sm, cfg)
