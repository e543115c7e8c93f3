"""DIM (Xie et al., CVPR 2019): with probability ``diversity_prob`` the input is bilinearly resized to a random size in
[S, int(S*resize_rate)), zero-padded at a random offset to int(S*resize_rate) and resized back to S.
Reference: transferattack/input_transformation/dim.py:35-68 (same constructor; the coin and the three integers are
drawn from torch's global CPU generator with the same calls in the same order, so a seeded run takes the same
decisions as the reference).

The three image-sized ops are ONE gather kernel (``ta_dim_fwd``: bulk-TMA-staged source rows, the intermediate
resize kept in shared memory) with a deterministic gather-form adjoint (``ta_dim_bwd``)."""
from ..utils import *
from .. import ops
from ..gradient.mifgsm import MIFGSM


class DIM(MIFGSM):
    graph_safe = False      # transform draws host-generator numbers on every call → never captured into a CUDA graph

    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, decay=1., resize_rate=1.1, diversity_prob=0.5, targeted=False,
                 random_start=False, norm='linfty', loss='crossentropy', device=None, attack='DIM', **kwargs):
        super().__init__(model_name, epsilon, alpha, epoch, decay, targeted, random_start, norm, loss, device, attack)
        if resize_rate < 1:
            raise Exception("Error! The resize rate should be larger than 1.")
        self.resize_rate = resize_rate
        self.diversity_prob = diversity_prob

    def draw(self, img_size):
        """The reference's CPU-generator draws (dim.py:47,54,60,62): coin, rnd, pad_top, pad_left.
        Returns None when the coin says 'do not transform', else (rnd, img_resize, top, left)."""
        if torch.rand(1) > self.diversity_prob:
            return None
        img_resize = int(img_size * self.resize_rate)
        rnd = torch.randint(low=min(img_size, img_resize), high=max(img_size, img_resize), size=(1,), dtype=torch.int32)
        rem = img_resize - rnd
        top = torch.randint(low=0, high=rem.item(), size=(1,), dtype=torch.int32)
        left = torch.randint(low=0, high=rem.item(), size=(1,), dtype=torch.int32)
        return int(rnd), img_resize, int(top), int(left)

    def transform(self, x, **kwargs):
        params = self.draw(x.shape[-1])
        if params is None:
            return x
        return ops.dim_resize_pad(x, *params)
