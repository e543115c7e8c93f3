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


DIM_GEOMETRIES = [(32, 33, 35, 1, 2), (32, 32, 35, 0, 3), (32, 34, 35, 1, 0), (20, 21, 22, 0, 0), (40, 43, 44, 1, 1),
                  (24, 24, 48, 10, 20), (33, 36, 36, 0, 0), (48, 52, 52, 0, 0), (16, 30, 40, 5, 5)]


@pytest.mark.parametrize("S,rnd,R,top,left", DIM_GEOMETRIES)
def test_dim_direct_models(S, rnd, R, top, left):
    """host tap / inverse tables and band table; forward: zero row + zero column as the padding → bit-identical to orc_dim_fwd
    (blend 1); both adjoint forms: every destination written exactly once, equal to the fp64 scatter oracle to rounding"""
    rng = np.random.default_rng(S * 100 + rnd)
    x = rng.random((1, S, S)).astype(np.float32); g = rng.standard_normal((1, S, S)).astype(np.float32)
    assert bits_equal(KM.dim_fwd_model(x, rnd, R, top, left), oracle.dim_fwd(x[None], rnd, R, top, left, blend=1)[0])
    want = oracle.dim_bwd(g[None], rnd, R, top, left)[0]
    for model in (KM.dim_bwd_scatter_model, KM.dim_bwd_gather_model):
        got = model(g, rnd, R, top, left)
        assert not np.isnan(got).any()
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
