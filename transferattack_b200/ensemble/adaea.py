"""AdaEA (Chen et al., ICCV 2023): adaptive ensemble attack — per-member loss weights from adaptive gradient modulation (AGM) and
a per-pixel disparity-reduced filter (DRF) that zeroes the ensemble gradient where the members' gradients disagree.
Reference: transferattack/ensemble/adaea.py:10-150 (same constructor and defaults incl. ``random_start=True``, ``agm`` / ``drf`` /
``get_adv_example`` hooks, same loop; ``self.model`` must be an ``EnsembleModel`` — the attack indexes ``self.model.models[k]``).

Kernels: the whole DRF (K(K-1)/2 normalise + cosine-similarity chains, the row means, the threshold and ``grad * mask``;
adaea.py:115-136, 74-76, 82) is ONE ``ta_adaea_drf`` launch; ``get_momentum`` + ``update_delta`` go through the base hooks (one
``ta_fused_tail`` launch when they are not overridden). AGM is K^2 surrogate forwards (torch)."""
import torch.nn.functional as F

from ..utils import *
from .. import ops
from ..attack import Attack


class AdaEA(Attack):
    def __init__(self, model_name, epsilon=16 / 255, alpha=1.6 / 255, epoch=10, decay=1.0, targeted=False,
                 random_start=True, beta=10, threshold=-0.3, norm='linfty', loss='crossentropy', device=None, attack='AdaEA', **kwargs):
        super().__init__(attack, model_name, epsilon, targeted, random_start, norm, loss, device)
        self.alpha, self.epoch, self.decay = alpha, epoch, decay
        # adaea.py:41 takes len(model_name); the loaded ensemble is the authority when load_model was customised
        self.num_model = self.model.num_models if isinstance(self.model, EnsembleModel) else len(model_name)
        self.beta = beta
        self.threshold = threshold

    def get_adv_example(self, ori_data, adv_data, grad):
        """adaea.py:138-148: one signed step from the current adversarial image, projected to the eps-ball and to [0, 1]"""
        step = adv_data.detach() + grad.sign() * self.alpha
        delta = torch.clamp(step - ori_data.detach(), -self.epsilon, self.epsilon)
        return torch.clamp(ori_data.detach() + delta, max=1.0, min=0.0)

    def agm(self, ori_data, cur_adv, grad, label):
        """adaea.py:87-113: w_j = softmax_j( beta * sum_{i != j} CE(f_i(x_j)) / CE(f_i(x_i)) ), x_k = member k's own one-step example"""
        ce = nn.CrossEntropyLoss()
        K = self.num_model
        adv = [self.get_adv_example(ori_data=ori_data, adv_data=cur_adv, grad=grad[k]) for k in range(K)]
        own = [ce(self.model.models[k](adv[k]), label) for k in range(K)]
        w = torch.zeros(size=(K,), device=self.device)
        for j in range(K):
            for i in range(K):
                if i != j:
                    w[j] += ce(self.model.models[i](adv[j]), label) / own[i] * self.beta
        return torch.softmax(w, dim=0)

    def drf(self, grads, data_size):
        """adaea.py:115-136: the un-thresholded reduce map [B, 1, H, W] (``ta_adaea_drf``)"""
        _, mp = ops.backend().adaea_drf(grads, self.threshold, None, want_map=True)
        return mp.view(data_size[0], 1, data_size[-2], data_size[-1])

    def forward(self, data, label, **kwargs):
        data = self._to_device(data).contiguous()
        label = self._to_device(label)
        B, C, H, W = data.size()
        ce = nn.CrossEntropyLoss()
        K = self.num_model
        be = ops.backend()
        momentum = 0.
        delta = torch.zeros_like(data).to(self.device) + 0.001 * torch.randn(data.shape, device=self.device)     # adaea.py:61
        delta.requires_grad = True
        fused_filter = type(self).drf is AdaEA.drf
        for _ in range(self.epoch):
            x = delta + data
            outputs = [self.model.models[k](x) for k in range(K)]
            losses = [ce(outputs[k], label) for k in range(K)]
            grads = [torch.autograd.grad(losses[k], delta, retain_graph=True, create_graph=False)[0] for k in range(K)]
            w = self.agm(ori_data=data, cur_adv=data + delta, grad=grads, label=label)
            output = (torch.stack(outputs, dim=0) * w.view(K, 1, 1)).sum(dim=0)
            loss = ce(output, label)
            grad = torch.autograd.grad(loss.sum(dim=0), delta)[0]
            if fused_filter:
                grad, _ = be.adaea_drf(grads, self.threshold, grad)               # DRF + threshold + grad * mask: one launch
            else:
                cos_res = self.drf(grads, data_size=(B, C, H, W))
                cos_res[cos_res >= self.threshold] = 1.
                cos_res[cos_res < self.threshold] = 0.
                grad = grad * cos_res
            momentum = self.get_momentum(grad, momentum)
            delta = self.update_delta(delta, data, momentum, self.alpha)
        return delta.detach()
