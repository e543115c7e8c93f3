"""NI-FGSM (Lin et al., ICLR 2020): gradient taken at the Nesterov look-ahead point x + alpha*decay*momentum.
Reference: transferattack/gradient/nifgsm.py:31-39. The look-ahead is one ``ta_stage_add`` launch
(out = x + coef * momentum, coef = fp32(alpha*decay)) wrapped in an identity-backward autograd node."""
from ..utils import *
from .. import ops
from .mifgsm import MIFGSM


class NIFGSM(MIFGSM):
    graph_safe = True       # hooks defined here are deterministic device code → capturable (attack.py: _graph_ok)

    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, decay=1., targeted=False, random_start=False,
                 norm='linfty', loss='crossentropy', device=None, attack='NI-FGSM', **kwargs):
        super().__init__(model_name, epsilon, alpha, epoch, decay, targeted, random_start, norm, loss, device, attack)

    def transform(self, x, momentum, **kwargs):
        if not torch.is_tensor(momentum):          # first iteration: momentum is the Python 0 → x + 0
            return x
        return ops.look_ahead(x, momentum, self.alpha * self.decay)
