"""VMI-FGSM (Wang & He, CVPR 2021): momentum of (gradient + variance), where the variance is the mean gradient over
``num_neighbor`` uniformly perturbed copies minus the current gradient.
Reference: transferattack/gradient/vmifgsm.py:33-97 (same constructor, ``get_variance`` and loop order; the
neighbour noise is drawn with the same torch call so the device generator is consumed identically).

Kernels per neighbour: ``ta_neighbor_stage`` ((data+delta)+noise, identity backward) and ``ta_accumulate``;
once per iteration ``ta_variance_finalize`` (acc/N - g) and ONE ``ta_fused_tail`` launch for `grad + variance` → mean →
momentum → update_delta → next `data + delta` (vmifgsm.py:86-97) when ``get_momentum`` / ``update_delta`` are the base
hooks: the kernel takes the variance as its addend and writes delta' into a second buffer, because the reference evaluates
the neighbours at the OLD delta after the momentum update (vmifgsm.py:90-94). Otherwise ``ta_add`` and the public hooks."""
from ..utils import *
from .. import ops
from ..attack import Attack


class VMIFGSM(Attack):
    #: draw the neighbour noise inside the staging kernel (``ta_neighbor_stage_philox``: torch's own Philox stream reproduced
    #: bit for bit, generator advanced as ``uniform_`` would) instead of ``zeros_like().uniform_()`` + a read of it
    philox_noise = os.environ.get("TA_B200_PHILOX", "1") == "1"

    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, beta=1.5, num_neighbor=20, epoch=10, decay=1., targeted=False,
                 random_start=False, norm='linfty', loss='crossentropy', device=None, attack='VMI-FGSM', **kwargs):
        super().__init__(attack, model_name, epsilon, targeted, random_start, norm, loss, device)
        self.alpha, self.epoch, self.decay = alpha, epoch, decay
        self.radius = beta * epsilon
        self.num_neighbor = num_neighbor

    def _neighbor_sum_batched(self, data, delta, label, momentum):
        """FAST MODE ONLY (Attack.fast_mode; not the parity path): the neighbours go through the surrogate several at a time.
        d/d delta of sum_k CE_k is the sum of the neighbour gradients, formed by autograd in one backward per chunk (its
        accumulation order, not the reference's running `grad +=`)."""
        B = data.shape[0]
        per = max(1, min(self.num_neighbor, self.fast_neighbor_images // max(B, 1)))
        acc = None
        for k0 in range(0, self.num_neighbor, per):
            n = min(per, self.num_neighbor - k0)
            xs = []
            for _ in range(n):
                if self.philox_noise and ops.philox_noise_available(delta):
                    xs.append(ops.neighbor_stage_philox(data, delta, -self.radius, self.radius))
                else:
                    noise = torch.zeros_like(delta).uniform_(-self.radius, self.radius).to(self.device)
                    xs.append(ops.neighbor_stage(data, delta, noise))
            x_all = torch.cat(xs, dim=0)
            mom = momentum.repeat(n, 1, 1, 1) if torch.is_tensor(momentum) else momentum
            loss = self.get_loss(self.get_logits(self.transform(x_all, momentum=mom)), label.repeat(n)) * n
            g = self.get_grad(loss, delta)
            acc = g if acc is None else acc + g
        return acc

    def get_variance(self, data, delta, label, cur_grad, momentum, **kwargs):
        be = ops.backend()
        if self.fast_mode and self.num_neighbor > 1:
            return be.variance_finalize(self._neighbor_sum_batched(data, delta, label, momentum), cur_grad, self.num_neighbor)
        acc = None
        for k in range(self.num_neighbor):
            if self.philox_noise and ops.philox_noise_available(delta):
                x_near = ops.neighbor_stage_philox(data, delta, -self.radius, self.radius)
            else:
                noise = torch.zeros_like(delta).uniform_(-self.radius, self.radius).to(self.device)
                x_near = ops.neighbor_stage(data, delta, noise)
            loss = self.get_loss(self.get_logits(self.transform(x_near, momentum=momentum)), label)
            acc = be.accumulate(acc, self.get_grad(loss, delta), first=(k == 0))
        return be.variance_finalize(acc, cur_grad, self.num_neighbor)

    def forward(self, data, label, **kwargs):
        if self.targeted:
            assert len(label) == 2
            label = label[1]
        data = self._to_device(data).contiguous()
        label = self._to_device(label)
        be = ops.backend()
        delta = self.init_delta(data)
        if self._fusable():
            return self._forward_fused(be, data, label, delta)
        momentum, variance = 0, None
        for _ in range(self.epoch):
            loss = self.get_loss(self.get_logits(self.transform(ops.stage_add(data, delta), momentum=momentum)), label)
            grad = self.get_grad(loss, delta)
            momentum = self.get_momentum(grad if variance is None else be.add(grad, variance), momentum)
            variance = self.get_variance(data, delta, label, grad, momentum)
            delta = self.update_delta(delta, data, momentum, self.alpha)
        return delta.detach()

    def _forward_fused(self, be, data, label, delta):
        """The same loop with the tail of every iteration as one launch. momentum is updated in place; delta ping-pongs between
        two leaves so that get_variance still sees the point the gradient was taken at."""
        kmode = self._mean_kernel_mode(data)
        m_buf = torch.empty_like(data)
        xadv = torch.empty_like(data)
        scale_out = torch.empty(data.shape[0], device=data.device, dtype=torch.float32)
        nxt = torch.empty_like(data).requires_grad_(True)
        momentum, variance, pre_x = None, None, None
        for _ in range(self.epoch):
            x = ops.stage_add(data, delta, precomputed=pre_x)
            loss = self.get_loss(self.get_logits(self.transform(x, momentum=0 if momentum is None else momentum)), label)
            grad = self.get_grad(loss, delta)
            self._tail(be, grad, momentum, m_buf, delta, nxt, data, xadv, scale_out, kmode, None, addend=variance)
            momentum, pre_x = m_buf, xadv
            variance = self.get_variance(data, delta, label, grad, momentum)
            delta, nxt = nxt, delta
        return delta.detach()
