"""SIM (Lin et al., ICLR 2020): the loss is averaged over ``num_scale`` copies x / 2^i of the input.
Reference: transferattack/input_transformation/sim.py:32-46. The replication is one ``ta_sim_fwd`` launch (x read once,
S scaled copies written) and its adjoint one ``ta_sim_bwd`` (S slices read, summed in autograd's order)."""
from ..utils import *
from .. import ops
from ..gradient.mifgsm import MIFGSM


class SIM(MIFGSM):
    graph_safe = True       # hooks defined here are deterministic device code → capturable (attack.py: _graph_ok)

    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, decay=1., num_scale=5, targeted=False, random_start=False,
                 norm='linfty', loss='crossentropy', device=None, attack='SIM', **kwargs):
        super().__init__(model_name, epsilon, alpha, epoch, decay, targeted, random_start, norm, loss, device, attack)
        self.num_scale = num_scale

    def transform(self, x, **kwargs):
        return ops.sim_scale(x, self.num_scale)

    def get_loss(self, logits, label):
        rep = label.repeat(self.num_scale)
        return -self.loss(logits, rep) if self.targeted else self.loss(logits, rep)
