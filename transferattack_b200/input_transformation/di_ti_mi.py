"""DI-TI-MI-FGSM: input diversity (DIM) on the way in, translation-invariant smoothing (TIM) on the gradient, momentum
update — the composite of BASELINE config 3. The reference tree has no such class (``class X(DIM, TIM)`` fails on its
positional constructor chain); it composes the two inline in advanced_objective/logit.py:66-99. Here it
is DIM's ``transform`` + TIM's ``get_grad`` on one MI-FGSM loop."""
from ..utils import *
from .. import ops
from .dim import DIM
from .tim import TIM


class DITIMI(DIM):
    graph_safe = True       # DIM's pre-drawn tables + TIM's deterministic convolution → capturable (see dim.py)
    conv_mode = TIM.conv_mode

    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, decay=1., resize_rate=1.1, diversity_prob=0.5,
                 kernel_type='gaussian', kernel_size=15, targeted=False, random_start=False, norm='linfty', loss='crossentropy',
                 device=None, attack='DI-TI-MI-FGSM', **kwargs):
        super().__init__(model_name, epsilon, alpha, epoch, decay, resize_rate, diversity_prob, targeted, random_start, norm, loss,
                         device, attack)
        self.kernel = self.generate_kernel(kernel_type, kernel_size)

    generate_kernel = TIM.generate_kernel
    smooth = TIM.smooth
    get_grad = TIM.get_grad


class SIDITIMI(DITIMI):
    """BASELINE config 3's "+SIM S=5" variant: the scale copies of SIM (sim.py:36-46) are formed first, then ONE
    DIM draw resizes / pads the whole S*B batch (dim.py:42-68 uses one (rnd, top, left) per call), TIM smooths the gradient —
    exactly what composing the reference's own hooks gives (``DIM.transform(SIM.transform(x))``, SIM's ``get_loss``,
    TIM's ``get_grad``). Kernels: ``ta_sim_fwd`` → ``ta_dim_fwd`` on S*B*3 planes, adjoints in reverse, ``ta_dwconv2d_sep``."""

    graph_safe = True

    def __init__(self, model_name, epsilon=16/255, alpha=1.6/255, epoch=10, decay=1., resize_rate=1.1, diversity_prob=0.5,
                 kernel_type='gaussian', kernel_size=15, num_scale=5, targeted=False, random_start=False, norm='linfty',
                 loss='crossentropy', device=None, attack='SI-DI-TI-MI-FGSM', **kwargs):
        super().__init__(model_name, epsilon, alpha, epoch, decay, resize_rate, diversity_prob, kernel_type, kernel_size, targeted,
                         random_start, norm, loss, device, attack)
        self.num_scale = num_scale

    def transform(self, x, **kwargs):
        return DIM.transform(self, ops.sim_scale(x, self.num_scale))

    def get_loss(self, logits, label):
        rep = label.repeat(self.num_scale)
        return -self.loss(logits, rep) if self.targeted else self.loss(logits, rep)
