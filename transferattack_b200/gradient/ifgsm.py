"""I-FGSM / BIM (Kurakin et al., ICLR-W 2017): the base loop with decay = 0, i.e. the momentum buffer is just the
L1-normalised gradient of the current iteration. Reference: transferattack/gradient/ifgsm.py:30-35."""
from ..utils import *
from ..attack import Attack


class IFGSM(Attack):
    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, targeted=False,
                 random_start=False, norm='linfty', loss='crossentropy', device=None, attack='I-FGSM', **kwargs):
        Attack.__init__(self, attack, model_name, epsilon, targeted, random_start, norm, loss, device)
        self.alpha, self.epoch, self.decay = alpha, epoch, 0     # decay 0: the buffer is just g / mean|g|
