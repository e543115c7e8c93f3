"""FGSM (Goodfellow et al., ICLR 2015): one iteration, step = epsilon, no momentum.
Reference: transferattack/gradient/fgsm.py:28-33."""
from ..utils import *
from ..attack import Attack


class FGSM(Attack):
    def __init__(self, model_name, epsilon=16/255, targeted=False, random_start=False,
                 norm='linfty', loss='crossentropy', device=None, **kwargs):
        Attack.__init__(self, 'FGSM', model_name, epsilon, targeted, random_start, norm, loss, device)
        self.alpha, self.epoch, self.decay = epsilon, 1, 0      # one step of size epsilon, no momentum
