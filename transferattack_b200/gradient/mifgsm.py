"""MI-FGSM (Dong et al., CVPR 2018) — configuration of the base loop: step size, iteration count, momentum decay.
Reference: transferattack/gradient/mifgsm.py:31-36 (same constructor signature and defaults:
epsilon=16/255, alpha=1.6/255, epoch=10, decay=1). All arithmetic runs in the base class's kernels; with no hook
overridden the whole tail of an iteration is one ``ta_fused_update_linf`` launch."""
from ..utils import *
from ..attack import Attack


class MIFGSM(Attack):
    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, decay=1., targeted=False,
                 random_start=False, norm='linfty', loss='crossentropy', device=None, attack='MI-FGSM', **kwargs):
        Attack.__init__(self, attack, model_name, epsilon, targeted, random_start, norm, loss, device)
        self.alpha, self.epoch, self.decay = alpha, epoch, decay
