"""TIM (Dong et al., CVPR 2019): the input gradient is smoothed with a fixed depthwise kernel before the momentum update.
Reference: transferattack/input_transformation/tim.py:37-73 (same constructor, same float64 kernel recipe for
gaussian / uniform / linear, ``self.kernel`` is the same [3,1,k,k] fp32 tensor).

``get_grad`` runs ``ta_dwconv2d_sep`` when ``self.kernel`` is still the generated (rank-1) kernel — row pass + column
pass in one CTA, 2k instead of k*k FMAs per element, HBM-bound — and ``ta_dwconv2d`` (direct k x k) otherwise or when
``conv_mode='direct'``.

Numerical contract (stated, not bit-parity): the reference convolves with the fp32 2-D kernel through ``F.conv2d`` (cuDNN /
ATen pick the summation order); the separable form multiplies by fp32(k1/sqrt(sum)) twice and sums 15 + 15 terms in tap order,
the direct form sums the 225 products in row-major tap order. Either way the smoothed gradient is within 1e-6 (relative to its
largest entry) of ``F.conv2d``'s (tests/test_kernels_gpu.py, golden tim.npz); the perturbation can therefore differ from the
reference's only where a momentum entry is zero to rounding (a sign tie) — measured: 0 differing elements over 10 iterations at
ResNet-50 B = 32 (profiles/e2e_parity_baseline_r2.json), asserted against the reference's own run-to-run floor."""
import numpy as np
import scipy.stats as st

from ..utils import *
from .. import ops
from ..gradient.mifgsm import MIFGSM


def make_kernel(kernel_type, kernel_size, nsig=3):
    """tim.py:42-66 in float64, cast to float32 at the end. Returns (K[3,1,k,k] f32, kcol[k] f32, krow[k] f32) with
    outer(kcol, krow) == K up to fp32 rounding (all three generated kernels are rank-1)."""
    kt = kernel_type.lower()
    if kt == 'gaussian':
        k1 = st.norm.pdf(np.linspace(-nsig, nsig, kernel_size))
    elif kt == 'uniform':
        k1 = np.ones(kernel_size)
    elif kt == 'linear':
        k1 = 1 - np.abs(np.linspace((-kernel_size+1)//2, (kernel_size-1)//2, kernel_size)/(kernel_size**2))
    else:
        raise Exception("Unspported kernel type {}".format(kernel_type))
    raw = np.outer(k1, k1)
    kernel = np.ones((kernel_size, kernel_size)) / (kernel_size ** 2) if kt == 'uniform' else raw / raw.sum()
    k2d = np.expand_dims(np.stack([kernel, kernel, kernel]), 1).astype(np.float32)
    factor = (k1 / np.sqrt(raw.sum())).astype(np.float32)
    return k2d, factor, factor.copy()


class TIM(MIFGSM):
    graph_safe = True       # hooks defined here are deterministic device code → capturable (attack.py: _graph_ok)

    conv_mode = 'separable'     # 'separable' | 'direct'

    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, decay=1., kernel_type='gaussian', kernel_size=15, targeted=False,
                 random_start=False, norm='linfty', loss='crossentropy', device=None, attack='TIM', **kwargs):
        super().__init__(model_name, epsilon, alpha, epoch, decay, targeted, random_start, norm, loss, device, attack)
        self.kernel = self.generate_kernel(kernel_type, kernel_size)

    def generate_kernel(self, kernel_type, kernel_size, nsig=3):
        k2d, kcol, krow = make_kernel(kernel_type, kernel_size, nsig)
        kernel = torch.from_numpy(k2d).to(self.device)
        hc, hr = np.stack([kcol] * 3), np.stack([krow] * 3)
        self._sep = (kernel, torch.from_numpy(hc).to(self.device), torch.from_numpy(hr).to(self.device), (hc, hr))
        return kernel

    def smooth(self, grad):
        be = ops.backend()
        sep = getattr(self, '_sep', None)
        if self.conv_mode == 'separable' and sep is not None and sep[0] is self.kernel and grad.shape[1] == 3:
            return be.dwconv2d_sep(grad, sep[1], sep[2], host=sep[3])
        return be.dwconv2d(grad, self.kernel.reshape(self.kernel.shape[0], self.kernel.shape[-2], self.kernel.shape[-1]))

    def get_grad(self, loss, delta, **kwargs):
        grad = torch.autograd.grad(loss, delta, retain_graph=False, create_graph=False)[0]
        return self.smooth(grad)
