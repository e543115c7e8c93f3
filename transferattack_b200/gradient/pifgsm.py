"""PI-FGSM (Gao et al., ECCV 2020): patch-wise iterative attack — the step is amplified by ``beta`` and the part of the
accumulated amplification that overflows the eps-ball is redistributed to the neighbourhood through a fixed depthwise
"project" kernel. Reference: transferattack/gradient/pifgsm.py:33-102 (same constructor and defaults, ``project_kern`` /
``project_noise`` / ``update_delta`` hooks with the same signatures, same loop order).

Per iteration, around the surrogate: ``ta_pi_cut_noise`` (amplification += beta*alpha*sign(m); cut noise), ``ta_dwconv2d``
(``project_noise``: 3x3 depthwise, weights 1/(k*k-1) with a zero centre), ``ta_pi_update_linf`` (projection = gamma*sign(conv),
amplification += projection, the L-inf step with the projection inside the eps-clip, box clamp) — 3 launches instead of the
reference's 16 ATen launches; ``get_momentum`` is the base hook. The L2 variant (pifgsm.py:64-66) keeps the reference's op
sequence in torch (no configuration uses it)."""
from ..utils import *
from .. import ops
from ..attack import Attack


class PIFGSM(Attack):
    def __init__(self, model_name, epsilon=16.0/255, alpha=1.6/255, epoch=10, decay=0., kern_size=3, gamma=16.0, beta=10.0,
                 targeted=False, random_start=False, norm='linfty', loss='crossentropy', device=None, attack='PI-FGSM', **kwargs):
        super().__init__(attack, model_name, epsilon, targeted, random_start, norm, loss, device)
        self.alpha = alpha
        self.epoch = epoch
        self.decay = decay
        self.kern_size = kern_size
        self.gamma = gamma / 255.0
        self.beta = beta

    def project_kern(self, kern_size):
        """pifgsm.py:45-52: ones/(k*k-1) with a zero centre, stacked [3,1,k,k]; returns (kernel on the device, padding)"""
        kern = np.ones((kern_size, kern_size), dtype=np.float32) / (kern_size ** 2 - 1)
        kern[kern_size // 2, kern_size // 2] = 0.0
        stack_kern = np.expand_dims(np.stack([kern.astype(np.float32)] * 3), 1)
        return torch.tensor(stack_kern).to(self.device), kern_size // 2

    def project_noise(self, x, stack_kern, padding_size):
        """pifgsm.py:55-58: F.conv2d(x, stack_kern, padding, groups=3) → ``ta_dwconv2d`` (odd kernel, 'same' zero padding)"""
        k = stack_kern.reshape(stack_kern.shape[0], stack_kern.shape[-2], stack_kern.shape[-1])
        if 2 * padding_size + 1 != k.shape[-1] or x.shape[1] != k.shape[0]:
            raise ValueError("project_noise: expects the [C,1,k,k] kernel and padding k//2 that project_kern returns")
        return ops.backend().dwconv2d(x, k)

    def update_delta(self, delta, data, grad, alpha, projection, **kwargs):
        """pifgsm.py:61-68 with a precomputed ``projection`` tensor (the public hook; the loop below uses the fused kernel)"""
        if self.norm == 'linfty':
            delta = torch.clamp(delta + alpha * grad.sign() + projection, -self.epsilon, self.epsilon)
        else:
            grad_norm = torch.norm(grad.view(grad.size(0), -1), dim=1).view(-1, 1, 1, 1)
            scaled_grad = grad / (grad_norm + 1e-20)
            delta = (delta + scaled_grad * alpha + projection).view(delta.size(0), -1).renorm(p=2, dim=0, maxnorm=self.epsilon).view_as(delta)
        delta = clamp(delta, img_min - data, img_max - data)
        return delta

    def forward(self, data, label, **kwargs):
        if self.targeted:
            assert len(label) == 2
            label = label[1]
        data = self._to_device(data).contiguous()
        label = self._to_device(label)
        be = ops.backend()
        delta = self.init_delta(data)
        stack_kern, padding_size = self.project_kern(self.kern_size)
        step = self.beta * self.alpha
        momentum, amplification = 0.0, None
        for _ in range(self.epoch):
            logits = self.get_logits(self.transform(ops.stage_add(data, delta)))
            loss = self.get_loss(logits, label)
            grad = self.get_grad(loss, delta)
            momentum = self.get_momentum(grad, momentum)
            with torch.no_grad():
                amplification, cut_noise = be.pi_cut_noise(amplification, momentum, step, self.epsilon)
                conv = self.project_noise(cut_noise, stack_kern, padding_size)
                if self.norm == 'linfty' and type(self).update_delta is PIFGSM.update_delta:
                    amplification, new_delta = be.pi_update_linf(delta, data, momentum, conv, amplification, step, self.gamma,
                                                                 self.epsilon, img_min, img_max)
                else:
                    projection = self.gamma * torch.sign(conv)
                    amplification = amplification + projection
                    new_delta = self.update_delta(delta.detach(), data, momentum, step, projection)
            delta = new_delta.detach().requires_grad_(True)
        return delta.detach()
