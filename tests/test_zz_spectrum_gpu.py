"""-m gpu: the tcgen05 spectrum transform (csrc/spectrum.cu) behind SSM (input_transformation/ssm.py:41-55):
four tensor-core GEMMs against the DCT-II matrix and its inverse. Oracle: the float64 matrix restatement
(oracle.spectrum_transform), itself pinned to the reference's FFT formulation in tests/test_reference_live.py.
Tolerance (floating point, stated): 3xTF32 → |out - f64| <= 2e-5 on [0,1] images (the reference's own fp32 FFT chain is within
1e-6 of float64; both are far below the transform's random jitter of eps = 0.063); single tf32 → <= 5e-3.
(Named test_zz_* so that it runs last: a fault in a tensor-core kernel would poison the CUDA context for later tests.)"""
import numpy as np
import pytest
import torch

import oracle
import transferattack_b200 as tab
from oracle import torch_ref
from helpers import make_attack, seed_all

# Opt-in: after the GPU call in which these tests first ran (all 6 passed on a B200; profiles/pytest_spectrum_r2.log) the box was
# reported unhealthy by the runner's post-call probe. The cause is not established (nothing in the run failed), so the
# tensor-core tests are kept out of the default `-m gpu` run until it is: TA_B200_TEST_TCGEN05=1 enables them.
import os
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TA_B200_TEST_TCGEN05", "0") != "1",
                                 reason="tcgen05 spectrum tests are opt-in (TA_B200_TEST_TCGEN05=1), see the comment above")]


@pytest.fixture(scope="module")
def be():
    from transferattack_b200 import ops
    ops._install_backend_for_tests(None)
    return ops.backend()


@pytest.mark.parametrize("B,N", [(2, 224), (1, 64), (3, 96), (64, 224)])
def test_spectrum_transform_matches_float64(be, B, N):
    g = torch.Generator().manual_seed(B * 1000 + N)
    x = torch.rand(B, 3, N, N, generator=g)
    gauss = torch.randn(B, 3, N, N, generator=g) * (16 / 255)
    mask = torch.rand(B, 3, N, N, generator=g) + 0.5
    n_ref = min(B, 4)
    ref = oracle.spectrum_transform(x[:n_ref].numpy(), gauss[:n_ref].numpy(), mask[:n_ref].numpy())
    out = be.spectrum_transform(x.cuda(), gauss.cuda(), mask.cuda(), 1).cpu().numpy()
    assert np.isfinite(out).all()
    err = np.abs(out[:n_ref] - ref).max()
    assert err <= 2e-5, err
    out0 = be.spectrum_transform(x.cuda(), gauss.cuda(), mask.cuda(), 0).cpu().numpy()
    assert np.abs(out0[:n_ref] - ref).max() <= 5e-3
    # no jitter: idct_2d(dct_2d(x)) == x
    ident = be.spectrum_transform(x.cuda(), None, None, 1).cpu()
    assert float((ident - x).abs().max()) <= 2e-5
    # against the reference's own FFT formulation on the GPU (restated in torch_ref.RefSSM, pinned live)
    r = torch_ref.RefSSM.__new__(torch_ref.RefSSM)
    fft = r.idct_2d(r.dct_2d(x.cuda() + gauss.cuda()) * mask.cuda()).cpu().numpy()
    assert np.abs(out - fft).max() <= 4e-5


def test_spectrum_transform_rejects_unsupported_sizes(be):
    with pytest.raises(RuntimeError):
        be.spectrum_transform(torch.zeros(1, 3, 20, 20, device="cuda"), None, None)
    with pytest.raises(ValueError):
        be.spectrum_transform(torch.zeros(1, 3, 32, 64, device="cuda"), None, None)


def test_ssm_native_runs_and_tracks_the_restatement():
    """Same random draws, same loop; the transform differs from the reference's fp32 FFT chain at the 1e-6 level, which a
    chaotic surrogate amplifies — so the check is statistical: valid perturbation, and most of it equal to the restatement's."""
    torch.manual_seed(0)
    import torchvision
    net = torchvision.models.resnet18(weights=None).eval().cuda()
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 224, 224, generator=g); y = torch.randint(0, 1000, (2,), generator=g)
    kw = {"num_spectrum": 3, "epoch": 3}
    seed_all(5); torch.cuda.manual_seed_all(5)
    ref = torch_ref.RefSSM(torch_ref.ref_wrap_model(net), **kw)(x, y)
    seed_all(5); torch.cuda.manual_seed_all(5)
    d = make_attack(tab, "ssm", net, **kw)(x, y)
    assert d.shape == ref.shape and float(d.abs().max()) <= 16 / 255 + 1e-7
    adv = x.cuda() + d
    assert float(adv.min()) >= 0 and float(adv.max()) <= 1
    agree = float((d == ref).float().mean())
    assert agree >= 0.9, agree
