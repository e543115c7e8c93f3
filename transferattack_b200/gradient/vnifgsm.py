"""VNI-FGSM: VMI-FGSM evaluated at the Nesterov look-ahead point (also for the neighbour samples).
Reference: transferattack/gradient/vnifgsm.py:33-41."""
from ..utils import *
from .. import ops
from .vmifgsm import VMIFGSM


class VNIFGSM(VMIFGSM):
    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, beta=1.5, num_neighbor=20, epoch=10, decay=1., targeted=False,
                 random_start=False, norm='linfty', loss='crossentropy', device=None, attack='VNI-FGSM', **kwargs):
        super().__init__(model_name, epsilon, alpha, beta, num_neighbor, epoch, decay, targeted, random_start, norm, loss, device, attack)

    def transform(self, x, momentum, **kwargs):
        if not torch.is_tensor(momentum):
            return x
        return ops.look_ahead(x, momentum, self.alpha * self.decay)
