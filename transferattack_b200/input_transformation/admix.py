"""Admix (Wang et al., ICCV 2021): each input is mixed with ``num_admix`` randomly chosen other images of the batch
(x + strength * x[perm]) and every mix is evaluated at ``num_scale`` scales.
Reference: transferattack/input_transformation/admix.py:34-51 (``torch.randperm`` on the CPU generator, one call per mix,
same order). One ``ta_admix_fwd`` launch writes all S*A*B images; ``ta_admix_bwd`` is the adjoint wrt the un-permuted
operand (the mixed-in image is detached in the reference). Mixing crosses samples, so this attack runs as
replicas, not batch shards, in multi-GPU mode."""
from ..utils import *
from .. import ops
from ..gradient.mifgsm import MIFGSM


class Admix(MIFGSM):
    graph_safe = False      # transform draws host-generator numbers on every call → never captured into a CUDA graph

    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, decay=1., num_scale=5, num_admix=3, admix_strength=0.2,
                 targeted=False, random_start=False, norm='linfty', loss='crossentropy', device=None, attack='Admix', **kwargs):
        super().__init__(model_name, epsilon, alpha, epoch, decay, targeted, random_start, norm, loss, device, attack)
        self.num_scale = num_scale
        self.num_admix = num_admix
        self.admix_strength = admix_strength

    def transform(self, x, **kwargs):
        perm = torch.stack([torch.randperm(x.size(0)) for _ in range(self.num_admix)]).to(device=x.device, dtype=torch.int32)
        return ops.admix_mix(x, perm.contiguous(), self.admix_strength, self.num_scale, self.num_admix)

    def get_loss(self, logits, label):
        rep = label.repeat(self.num_scale * self.num_admix)
        return -self.loss(logits, rep) if self.targeted else self.loss(logits, rep)
