"""EMI-FGSM (Wang et al., BMVC 2021): the gradient is averaged over ``num_sample`` points sampled along the previous
iteration's L1-normalised gradient, x + c_k * alpha * g_bar with c = linspace(-radius, radius, num_sample).
Reference: transferattack/gradient/emifgsm.py:33-105 (same constructor, factors, label repetition, loop order).

The K-way replication is one ``ta_lin_sample_fwd`` launch (reads x and g_bar once, writes K copies) and its adjoint
one ``ta_lin_sample_bwd`` (sums the K gradient slices in autograd's accumulation order). With the base ``get_momentum`` /
``update_delta`` the tail of an iteration — bar_grad = g / mean|g|, momentum, update_delta, next `data + delta`
(emifgsm.py:97-103) — is ONE ``ta_fused_tail`` launch that also emits bar_grad."""
from ..utils import *
from .. import ops
from .mifgsm import MIFGSM


class EMIFGSM(MIFGSM):
    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, decay=1., num_sample=11, radius=7, sample_method='linear',
                 targeted=False, random_start=False, norm='linfty', loss='crossentropy', device=None, attack='EMI-FGSM', **kwargs):
        super().__init__(model_name, epsilon, alpha, epoch, decay, targeted, random_start, norm, loss, device, attack)
        self.num_sample = num_sample
        self.radius = radius
        self.sample_method = sample_method.lower()

    def get_factors(self):
        """emifgsm.py:40-51 (linear / uniform / gaussian sampling of the factors)."""
        if self.sample_method == 'linear':
            return np.linspace(-self.radius, self.radius, num=self.num_sample)
        if self.sample_method == 'uniform':
            return np.random.uniform(-self.radius, self.radius, size=self.num_sample)
        if self.sample_method == 'gaussian':
            return np.clip(np.random.normal(size=self.num_sample)/3, -1, 1)*self.radius
        raise Exception('Unsupported sampling method {}!'.format(self.sample_method))

    def transform(self, x, grad, **kwargs):
        """emifgsm.py:53-58 always samples linearly, whatever ``sample_method`` says; so does this."""
        factors = np.linspace(-self.radius, self.radius, num=self.num_sample)
        coefs = [float(np.float32(f * self.alpha)) for f in factors]
        return ops.lin_sample(x, grad if torch.is_tensor(grad) else None, coefs)

    def get_loss(self, logits, label):
        rep = label.repeat(self.num_sample)
        return -self.loss(logits, rep) if self.targeted else self.loss(logits, rep)

    def forward(self, data, label, **kwargs):
        if self.targeted:
            assert len(label) == 2
            label = label[1]
        data = self._to_device(data).contiguous()
        label = self._to_device(label)
        be = ops.backend()
        delta = self.init_delta(data)
        if self._fusable():
            kmode = self._mean_kernel_mode(data)
            m_buf, xadv, bar_buf = torch.empty_like(data), torch.empty_like(data), torch.empty_like(data)
            scale_out = torch.empty(data.shape[0], device=data.device, dtype=torch.float32)
            momentum, bar_grad, pre_x = None, 0, None
            for _ in range(self.epoch):
                x = ops.stage_add(data, delta, precomputed=pre_x)
                loss = self.get_loss(self.get_logits(self.transform(x, grad=bar_grad)), label)
                grad = self.get_grad(loss, delta)
                self._tail(be, grad, momentum, m_buf, delta, delta, data, xadv, scale_out, kmode, None, gbar_out=bar_buf)
                momentum, bar_grad, pre_x = m_buf, bar_buf, xadv
            return delta.detach()
        momentum, bar_grad = 0, 0
        for _ in range(self.epoch):
            loss = self.get_loss(self.get_logits(self.transform(ops.stage_add(data, delta), grad=bar_grad)), label)
            grad = self.get_grad(loss, delta)
            bar_grad = be.momentum(grad, None, self._abs_mean(grad), 0.0)      # grad / mean|grad|  (emifgsm.py:97)
            momentum = self.get_momentum(grad, momentum)
            delta = self.update_delta(delta, data, momentum, self.alpha)
        return delta.detach()
