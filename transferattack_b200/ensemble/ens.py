"""ENS (Liu et al., ICLR 2017): MI-FGSM on the mean logits of several surrogates (``model_name`` is a list, the base's
``load_model`` wraps the members in ``EnsembleModel``). Reference: transferattack/ensemble/ens.py:31-36.
For one-surrogate-per-GPU execution see ``transferattack_b200.multigpu.ShardedEnsembleModel``."""
from ..utils import *
from ..attack import Attack


class ENS(Attack):
    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, decay=1., targeted=False,
                 random_start=False, norm='linfty', loss='crossentropy', device=None, attack='ENS', **kwargs):
        Attack.__init__(self, attack, model_name, epsilon, targeted, random_start, norm, loss, device)
        self.alpha, self.epoch, self.decay = alpha, epoch, decay
