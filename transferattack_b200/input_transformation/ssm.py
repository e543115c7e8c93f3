"""SSM — Spectrum Simulation Attack (Long et al., ECCV 2022): the gradient is averaged over ``num_spectrum`` copies of the input
whose DCT spectrum is jittered: x_idct = idct_2d(dct_2d(x + gauss) * mask), gauss ~ N(0, eps^2), mask ~ U(1-rho, 1+rho).
Reference: transferattack/input_transformation/ssm.py:8-200 (same constructor and defaults, same random draws in the same
order — ``torch.randn`` on the HOST generator for gauss, ``rand_like`` on the device generator for the mask —, same loop: the
gradient is taken with respect to x_idct itself, ssm.py:88, so nothing is differentiated through the transform).

The reference evaluates the 224-point DCT-II and its inverse through FFTs (about 40 ATen launches per transform). Here the
whole transform is ``ta_spectrum_transform``: four tensor-core GEMMs against the constant DCT matrix and its inverse
(tcgen05, 3xTF32 operands, fp32 accumulation in TMEM; csrc/spectrum.cu), with the ``x + gauss`` add and the mask product
fused into their load / epilogue. The result agrees with the float64 transform to ~1e-5 (tests); the reference's own fp32 FFT
chain deviates from float64 by a similar amount, so attack-level equality is statistical, not bitwise (the transform is
randomised by construction)."""
from ..utils import *
from .. import ops
from ..gradient.mifgsm import MIFGSM


class SSM(MIFGSM):
    #: 1 = 3xTF32 (fp32-level products), 0 = single tf32 product per term (faster, ~1e-3 relative)
    spectrum_precision = 1

    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, decay=1., num_spectrum=20, rho=0.5, targeted=False,
                 random_start=False, norm='linfty', loss='crossentropy', device=None, **kwargs):
        super().__init__(model_name, epsilon, alpha, epoch, decay, targeted, random_start, norm, loss, device)
        self.num_spectrum = num_spectrum
        self.epsilon = epsilon
        self.rho = rho

    def transform(self, x, **kwargs):
        """ssm.py:41-55. The two random tensors are drawn exactly as the reference draws them."""
        gauss = (torch.randn(x.size()[0], 3, 224, 224) * self.epsilon).to(x.device)
        mask = torch.rand_like(x) * 2 * self.rho + 1 - self.rho
        return ops.backend().spectrum_transform(x, gauss, mask, self.spectrum_precision)

    def forward(self, data, label, **kwargs):
        if self.targeted:
            assert len(label) == 2
            label = label[1]
        data = self._to_device(data).contiguous()
        label = self._to_device(label)
        be = ops.backend()
        delta = self.init_delta(data)
        momentum = 0
        for _ in range(self.epoch):
            grads = None
            for k in range(self.num_spectrum):
                with torch.no_grad():
                    x_idct = self.transform(ops.stage_add(data, delta.detach()))
                x_idct = x_idct.detach().requires_grad_(True)            # the reference differentiates w.r.t. x_idct (ssm.py:88)
                loss = self.get_loss(self.get_logits(x_idct), label)
                grads = be.accumulate(grads, self.get_grad(loss, x_idct), first=(k == 0))
            grads = grads / self.num_spectrum
            momentum = self.get_momentum(grads, momentum)
            delta = self.update_delta(delta, data, momentum, self.alpha)
        return delta.detach()
