"""GRA (Zhu et al., ICCV 2023): the gradient is blended with the average gradient of ``num_neighbor`` uniformly perturbed copies by
their cosine similarity, and a per-element decay indicator M shrinks the step wherever the momentum's sign flipped.
Reference: transferattack/gradient/gra.py:33-153 (same constructor and defaults, ``get_average_gradient`` /
``get_cosine_similarity`` / ``get_decay_indicator`` hooks with the same signatures, same loop order, eta = 0.94).

Kernels: per neighbour ``ta_neighbor_stage(_philox)`` (the reference's uniform_ stream, drawn in the kernel) and ``ta_accumulate``;
per iteration ONE ``ta_gra_update`` launch for the decay indicator + the tensor-step ``update_delta`` (the reference: 17
elementwise launches). The cosine similarity (three per-sample sums) and the blend stay torch ops: their summation order is
torch's, which keeps the result bit-identical to the reference."""
from ..utils import *
from .. import ops
from ..attack import Attack


class GRA(Attack):
    philox_noise = os.environ.get("TA_B200_PHILOX", "1") == "1"

    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, beta=3.5, num_neighbor=20, epoch=10, decay=1., targeted=False,
                 random_start=False, norm='linfty', loss='crossentropy', device=None, attack='GRA', **kwargs):
        super().__init__(attack, model_name, epsilon, targeted, random_start, norm, loss, device)
        self.alpha, self.epoch, self.decay = alpha, epoch, decay
        self.radius = beta * epsilon
        self.num_neighbor = num_neighbor

    def get_average_gradient(self, data, delta, label, momentum, **kwargs):
        """gra.py:42-58: mean of the gradients at `num_neighbor` points data + delta + U(-radius, radius)"""
        be = ops.backend()
        acc = None
        for k in range(self.num_neighbor):
            if self.philox_noise and ops.philox_noise_available(delta):
                x_near = ops.neighbor_stage_philox(data, delta, -self.radius, self.radius)
            else:
                noise = torch.zeros_like(delta).uniform_(-self.radius, self.radius).to(self.device)
                x_near = ops.neighbor_stage(data, delta, noise)
            loss = self.get_loss(self.get_logits(self.transform(x_near, momentum=momentum)), label)
            acc = be.accumulate(acc, self.get_grad(loss, delta), first=(k == 0))
        return acc / self.num_neighbor

    def get_cosine_similarity(self, cur_grad, sam_grad, **kwargs):
        """gra.py:60-72 (per-sample; torch's own reductions, so the same bits as the reference)"""
        cur = cur_grad.view(cur_grad.size(0), -1)
        sam = sam_grad.view(sam_grad.size(0), -1)
        cos = torch.sum(cur * sam, dim=1) / (torch.sqrt(torch.sum(cur ** 2, dim=1)) * torch.sqrt(torch.sum(sam ** 2, dim=1)))
        return cos.unsqueeze(-1).unsqueeze(-1).unsqueeze(-1)

    def get_decay_indicator(self, M, delta, cur_noise, last_noise, eta, **kwargs):
        """gra.py:74-93 as a public hook: M * (eq + (1 - eq) * eta) — the ``ta_gra_update`` kernel with a zero step"""
        last = None if isinstance(last_noise, (int, float)) and last_noise == 0 else last_noise
        if last is not None and not torch.is_tensor(last):
            last = torch.full_like(cur_noise, float(last))
        M_new, _ = ops.backend().gra_update(M, last, cur_noise, eta, 0.0, delta, delta, self.epsilon, 0.0, 0.0)
        return M_new

    def forward(self, data, label, **kwargs):
        if self.targeted:
            assert len(label) == 2
            label = label[1]
        data = self._to_device(data).contiguous()
        label = self._to_device(label)
        be = ops.backend()
        delta = self.init_delta(data)
        eta = 0.94
        M = torch.full_like(delta, 1 / eta)
        cls = type(self)
        fused = (self.norm == 'linfty' and cls.get_decay_indicator is GRA.get_decay_indicator and cls.update_delta is Attack.update_delta
                 and isinstance(self.alpha, (int, float)))
        momentum = 0
        for _ in range(self.epoch):
            loss = self.get_loss(self.get_logits(self.transform(ops.stage_add(data, delta), momentum=momentum)), label)
            grad = self.get_grad(loss, delta)
            samgrad = self.get_average_gradient(data, delta, label, momentum)
            s = self.get_cosine_similarity(grad, samgrad)
            current_grad = s * grad + (1 - s) * samgrad
            last_momentum = momentum
            momentum = self.get_momentum(current_grad, momentum)
            if fused:
                last = None if not torch.is_tensor(last_momentum) else last_momentum
                M, d_new = be.gra_update(M, last, momentum, eta, self.alpha, delta, data, self.epsilon, img_min, img_max)
                delta = d_new.requires_grad_(True)
            else:
                M = self.get_decay_indicator(M, delta, momentum, last_momentum, eta)
                delta = self.update_delta(delta, data, momentum, M * self.alpha)
        return delta.detach()
