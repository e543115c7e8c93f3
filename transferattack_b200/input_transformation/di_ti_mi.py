"""DI-TI-MI-FGSM: input diversity (DIM) on the way in, translation-invariant smoothing (TIM) on the gradient, momentum
update — the composite of BASELINE config 3. The reference tree has no such class (``class X(DIM, TIM)`` fails on its
positional constructor chain); it composes the two inline in advanced_objective/logit.py:66-99. Here it
is DIM's ``transform`` + TIM's ``get_grad`` on one MI-FGSM loop."""
from ..utils import *
from .dim import DIM
from .tim import TIM


class DITIMI(DIM):
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
