"""TEST INFRASTRUCTURE ONLY — eager-PyTorch restatement of the reference hot loop.

This is the end-to-end comparator: it issues the same ATen ops, in the same order, as the reference's
``transferattack/attack.py`` and the in-scope plugins, so that on one device with one surrogate it
reproduces the reference's perturbation bit for bit (checked against the live reference in
``tests/test_reference_live.py`` whenever ``/root/reference`` is present, and against
``tests/golden/e2e_*.npz``).  ``bench.py`` times it on the host cores as the CPU baseline
(``cpu_baseline.kind == "port"``; the Python reference itself cannot travel to the GPU box).

Nothing under ``transferattack_b200/`` imports this module.

Citations: file:line under the reference's ``transferattack/`` directory.
"""
import numpy as np
import scipy.stats as st
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision.transforms as T


# ---- utils.py:37-79 -------------------------------------------------------------------------------
class RefPreprocess(nn.Module):
    """utils.py:72-79 — torchvision Resize then Normalize (clone, sub_, div_ + a host sync)."""

    def __init__(self, resize, mean, std):
        super().__init__()
        self.resize = T.Resize(resize)
        self.normalize = T.Normalize(mean, std)

    def forward(self, x):
        return self.normalize(self.resize(x))


def ref_wrap_model(model):
    """utils.py:37-60: timm default_cfg mean/std, Inception → 0.5/0.5 @299, else ImageNet @224."""
    if hasattr(model, "default_cfg"):
        mean, std, size = model.default_cfg["mean"], model.default_cfg["std"], 224
    elif "Inc" in model.__class__.__name__:
        mean, std, size = [0.5] * 3, [0.5] * 3, 299
    else:
        mean, std, size = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], 224
    return nn.Sequential(RefPreprocess(size, mean, std), model)


class RefEnsemble(nn.Module):
    """utils.py:82-105 — members run sequentially on one device; stack → mean(dim=0)."""

    def __init__(self, members, mode="mean"):
        super().__init__()
        self.device = next(members[0].parameters()).device
        self.models = [m.to(self.device) for m in members]
        self.num_models = len(members)
        self.mode = mode

    def forward(self, x):
        outs = torch.stack([m(x) for m in self.models], dim=0)
        if self.mode == "mean":
            return torch.mean(outs, dim=0)
        if self.mode == "ind":
            return outs
        raise NotImplementedError


def _box(x, lo, hi):
    """utils.py:68-69"""
    return torch.min(torch.max(x, lo), hi)


# ---- attack.py:8-169 --------------------------------------------------------------------------------
class RefAttack:
    """The reference loop with its eight hooks, eager ATen ops only (attack.py:67-153)."""

    def __init__(self, model, epsilon=16 / 255, alpha=1.6 / 255, epoch=10, decay=1.0, targeted=False,
                 random_start=False, norm="linfty", device=None):
        if norm not in ("l2", "linfty"):
            raise Exception("Unsupported norm {}".format(norm))
        self.model = model
        self.epsilon, self.alpha, self.epoch, self.decay = epsilon, alpha, epoch, decay
        self.targeted, self.random_start, self.norm = targeted, random_start, norm
        if isinstance(model, RefEnsemble):
            self.device = model.device
        else:
            self.device = next(model.parameters()).device if device is None else device
        self.loss = nn.CrossEntropyLoss()
        self.trace = None  # optional list collecting (grad, momentum, delta) per iteration

    # hooks -------------------------------------------------------------------------------------------
    def transform(self, x, **kw):
        return x

    def get_logits(self, x, **kw):
        return self.model(x)

    def get_loss(self, logits, label):
        v = self.loss(logits, label)
        return -v if self.targeted else v

    def get_grad(self, loss, delta, **kw):
        return torch.autograd.grad(loss, delta, retain_graph=False, create_graph=False)[0]

    def get_momentum(self, grad, momentum, **kw):
        return momentum * self.decay + grad / (grad.abs().mean(dim=(1, 2, 3), keepdim=True))

    def init_delta(self, data, **kw):
        delta = torch.zeros_like(data).to(self.device)
        if self.random_start:
            if self.norm == "linfty":
                delta.uniform_(-self.epsilon, self.epsilon)
            else:
                delta.normal_(-self.epsilon, self.epsilon)
                flat = delta.view(delta.size(0), -1)
                nrm = flat.norm(p=2, dim=-1).view(delta.size(0), 1, 1, 1)
                r = torch.zeros_like(data).uniform_(0, 1).to(self.device)
                delta *= r / nrm * self.epsilon
            delta = _box(delta, 0 - data, 1.0 - data)
        delta.requires_grad = True
        return delta

    def update_delta(self, delta, data, grad, alpha, **kw):
        if self.norm == "linfty":
            delta = torch.clamp(delta + alpha * grad.sign(), -self.epsilon, self.epsilon)
        else:
            gnorm = torch.norm(grad.view(grad.size(0), -1), dim=1).view(-1, 1, 1, 1)
            ghat = grad / (gnorm + 1e-20)
            delta = (delta + ghat * alpha).view(delta.size(0), -1).renorm(p=2, dim=0, maxnorm=self.epsilon).view_as(delta)
        delta = _box(delta, 0 - data, 1.0 - data)
        return delta.detach().requires_grad_(True)

    # loop --------------------------------------------------------------------------------------------
    def _prep(self, data, label):
        if self.targeted:
            assert len(label) == 2
            label = label[1]
        return data.clone().detach().to(self.device), label.clone().detach().to(self.device)

    def forward(self, data, label, **kw):
        data, label = self._prep(data, label)
        delta = self.init_delta(data)
        momentum = 0
        for _ in range(self.epoch):
            logits = self.get_logits(self.transform(data + delta, momentum=momentum))
            loss = self.get_loss(logits, label)
            grad = self.get_grad(loss, delta)
            momentum = self.get_momentum(grad, momentum)
            delta = self.update_delta(delta, data, momentum, self.alpha)
            if self.trace is not None:
                self.trace.append((grad.detach().clone(), momentum.detach().clone(), delta.detach().clone()))
        return delta.detach()

    def __call__(self, *a, **kw):
        self.model.eval()
        return self.forward(*a, **kw)


def ref_mifgsm(model, **kw):          # gradient/mifgsm.py:31-36
    return RefAttack(model, **kw)


def ref_ifgsm(model, **kw):           # gradient/ifgsm.py:30-35  (decay = 0)
    kw = dict(kw); kw["decay"] = 0
    return RefAttack(model, **kw)


def ref_fgsm(model, epsilon=16 / 255, **kw):  # gradient/fgsm.py:28-33
    return RefAttack(model, epsilon=epsilon, alpha=epsilon, epoch=1, decay=0, **kw)


class RefNIFGSM(RefAttack):           # gradient/nifgsm.py:35-39
    def transform(self, x, momentum, **kw):
        return x + self.alpha * self.decay * momentum


class RefDIM(RefAttack):              # input_transformation/dim.py:42-68
    def __init__(self, model, resize_rate=1.1, diversity_prob=0.5, **kw):
        super().__init__(model, **kw)
        if resize_rate < 1:
            raise Exception("Error! The resize rate should be larger than 1.")
        self.resize_rate, self.diversity_prob = resize_rate, diversity_prob
        self.last_params = None

    def transform(self, x, **kw):
        if torch.rand(1) > self.diversity_prob:
            self.last_params = None
            return x
        size = x.shape[-1]
        big = int(size * self.resize_rate)
        rnd = torch.randint(low=min(size, big), high=max(size, big), size=(1,), dtype=torch.int32)
        y1 = F.interpolate(x, size=[rnd, rnd], mode="bilinear", align_corners=False)
        rem = big - rnd
        top = torch.randint(low=0, high=rem.item(), size=(1,), dtype=torch.int32)
        left = torch.randint(low=0, high=rem.item(), size=(1,), dtype=torch.int32)
        self.last_params = (int(rnd), big, int(top), int(left))
        y2 = F.pad(y1, [left.item(), (rem - left).item(), top.item(), (rem - top).item()], value=0)
        return F.interpolate(y2, size=[size, size], mode="bilinear", align_corners=False)


def ref_tim_kernel(kernel_type="gaussian", kernel_size=15, nsig=3):
    """input_transformation/tim.py:42-66 — float64 numpy → float32 [3,1,k,k]."""
    kt = kernel_type.lower()
    if kt == "gaussian":
        k1 = st.norm.pdf(np.linspace(-nsig, nsig, kernel_size))
        raw = np.outer(k1, k1)
        k = raw / raw.sum()
    elif kt == "uniform":
        k = np.ones((kernel_size, kernel_size)) / (kernel_size ** 2)
    elif kt == "linear":
        k1 = 1 - np.abs(np.linspace((-kernel_size + 1) // 2, (kernel_size - 1) // 2, kernel_size) / (kernel_size ** 2))
        raw = np.outer(k1, k1)
        k = raw / raw.sum()
    else:
        raise Exception("Unspported kernel type {}".format(kernel_type))
    return torch.from_numpy(np.expand_dims(np.stack([k, k, k]), 1).astype(np.float32))


class RefTIM(RefAttack):              # input_transformation/tim.py:68-73
    def __init__(self, model, kernel_type="gaussian", kernel_size=15, **kw):
        super().__init__(model, **kw)
        self.kernel = ref_tim_kernel(kernel_type, kernel_size).to(self.device)

    def get_grad(self, loss, delta, **kw):
        g = torch.autograd.grad(loss, delta, retain_graph=False, create_graph=False)[0]
        return F.conv2d(g, self.kernel, stride=1, padding="same", groups=3)


class RefSIM(RefAttack):              # input_transformation/sim.py:36-46
    def __init__(self, model, num_scale=5, **kw):
        super().__init__(model, **kw)
        self.num_scale = num_scale

    def transform(self, x, **kw):
        return torch.cat([x / (2 ** i) for i in range(self.num_scale)])

    def get_loss(self, logits, label):
        v = self.loss(logits, label.repeat(self.num_scale))
        return -v if self.targeted else v


class RefAdmix(RefAttack):            # input_transformation/admix.py:40-51
    def __init__(self, model, num_scale=5, num_admix=3, admix_strength=0.2, **kw):
        super().__init__(model, **kw)
        self.num_scale, self.num_admix, self.admix_strength = num_scale, num_admix, admix_strength

    def transform(self, x, **kw):
        mixed = torch.concat([(x + self.admix_strength * x[torch.randperm(x.size(0))].detach())
                              for _ in range(self.num_admix)], dim=0)
        return torch.concat([mixed / (2 ** i) for i in range(self.num_scale)])

    def get_loss(self, logits, label):
        v = self.loss(logits, label.repeat(self.num_scale * self.num_admix))
        return -v if self.targeted else v


class RefDITIMI(RefDIM):
    """Config 3 composite (no such class in the reference tree): DIM's transform with
    TIM's get_grad, as advanced_objective/logit.py:66-99 composes them inline."""

    def __init__(self, model, kernel_type="gaussian", kernel_size=15, **kw):
        super().__init__(model, **kw)
        self.kernel = ref_tim_kernel(kernel_type, kernel_size).to(self.device)

    get_grad = RefTIM.get_grad


class RefSIDITIMI(RefDITIMI):
    """Config 3's "+SIM" variant composed from the reference's hooks: DIM.transform(SIM.transform(x)) (sim.py:36-46 then
    dim.py:42-68, one DIM draw for the S*B batch), SIM's get_loss, TIM's get_grad."""

    def __init__(self, model, num_scale=5, **kw):
        super().__init__(model, **kw)
        self.num_scale = num_scale

    def transform(self, x, **kw):
        return RefDIM.transform(self, torch.cat([x / (2 ** i) for i in range(self.num_scale)]))

    def get_loss(self, logits, label):
        v = self.loss(logits, label.repeat(self.num_scale))
        return -v if self.targeted else v


class RefVMIFGSM(RefAttack):          # gradient/vmifgsm.py:42-97
    def __init__(self, model, beta=1.5, num_neighbor=20, **kw):
        super().__init__(model, **kw)
        self.radius = beta * self.epsilon
        self.num_neighbor = num_neighbor

    def get_variance(self, data, delta, label, cur_grad, momentum, **kw):
        grad = 0
        for _ in range(self.num_neighbor):
            noise = torch.zeros_like(delta).uniform_(-self.radius, self.radius).to(self.device)
            logits = self.get_logits(self.transform(data + delta + noise, momentum=momentum))
            loss = self.get_loss(logits, label)
            grad += self.get_grad(loss, delta)
        return grad / self.num_neighbor - cur_grad

    def forward(self, data, label, **kw):
        data, label = self._prep(data, label)
        delta = self.init_delta(data)
        momentum, variance = 0, 0
        for _ in range(self.epoch):
            logits = self.get_logits(self.transform(data + delta, momentum=momentum))
            loss = self.get_loss(logits, label)
            grad = self.get_grad(loss, delta)
            momentum = self.get_momentum(grad + variance, momentum)
            variance = self.get_variance(data, delta, label, grad, momentum)
            delta = self.update_delta(delta, data, momentum, self.alpha)
        return delta.detach()


class RefVNIFGSM(RefVMIFGSM):         # gradient/vnifgsm.py:37-41
    def transform(self, x, momentum, **kw):
        return x + self.alpha * self.decay * momentum


class RefEMIFGSM(RefAttack):          # gradient/emifgsm.py:53-105
    def __init__(self, model, num_sample=11, radius=7, **kw):
        super().__init__(model, **kw)
        self.num_sample, self.radius = num_sample, radius

    def transform(self, x, grad, **kw):
        factors = np.linspace(-self.radius, self.radius, num=self.num_sample)
        return torch.concat([x + f * self.alpha * grad for f in factors])

    def get_loss(self, logits, label):
        v = self.loss(logits, label.repeat(self.num_sample))
        return -v if self.targeted else v

    def forward(self, data, label, **kw):
        data, label = self._prep(data, label)
        delta = self.init_delta(data)
        momentum, bar_grad = 0, 0
        for _ in range(self.epoch):
            logits = self.get_logits(self.transform(data + delta, grad=bar_grad))
            loss = self.get_loss(logits, label)
            grad = self.get_grad(loss, delta)
            bar_grad = grad / (grad.abs().mean(dim=(1, 2, 3), keepdim=True))
            momentum = self.get_momentum(grad, momentum)
            delta = self.update_delta(delta, data, momentum, self.alpha)
        return delta.detach()


class RefPIFGSM(RefAttack):           # gradient/pifgsm.py:33-102 (device-agnostic: the reference hard-codes .cuda())
    def __init__(self, model, epsilon=16.0 / 255, alpha=1.6 / 255, epoch=10, decay=0., kern_size=3, gamma=16.0, beta=10.0, **kw):
        super().__init__(model, epsilon=epsilon, alpha=alpha, epoch=epoch, decay=decay, **kw)
        self.kern_size, self.gamma, self.beta = kern_size, gamma / 255.0, beta

    def project_kern(self, kern_size):
        kern = np.ones((kern_size, kern_size), dtype=np.float32) / (kern_size ** 2 - 1)
        kern[kern_size // 2, kern_size // 2] = 0.0
        kern = kern.astype(np.float32)
        stack_kern = np.expand_dims(np.stack([kern, kern, kern]), 1)
        return torch.tensor(stack_kern).to(self.device), kern_size // 2

    def project_noise(self, x, stack_kern, padding_size):
        return F.conv2d(x, stack_kern, padding=(padding_size, padding_size), groups=3)

    def update_delta(self, delta, data, grad, alpha, projection, **kw):
        if self.norm == "linfty":
            delta = torch.clamp(delta + alpha * grad.sign() + projection, -self.epsilon, self.epsilon)
        else:
            gnorm = torch.norm(grad.view(grad.size(0), -1), dim=1).view(-1, 1, 1, 1)
            delta = (delta + grad / (gnorm + 1e-20) * alpha + projection).view(delta.size(0), -1).renorm(p=2, dim=0, maxnorm=self.epsilon).view_as(delta)
        return _box(delta, 0 - data, 1.0 - data)

    def forward(self, data, label, **kw):
        data, label = self._prep(data, label)
        delta = self.init_delta(data)
        delta.requires_grad = True
        stack_kern, padding_size = self.project_kern(self.kern_size)
        momentum, amplification = 0.0, 0.0
        for _ in range(self.epoch):
            logits = self.get_logits(self.transform(data + delta))
            loss = self.get_loss(logits, label)
            grad = self.get_grad(loss, delta)
            momentum = self.get_momentum(grad, momentum)
            amplification += self.beta * self.alpha * momentum.sign()
            cut_noise = torch.clamp(abs(amplification) - self.epsilon, 0, 10000.0) * torch.sign(amplification)
            projection = self.gamma * torch.sign(self.project_noise(cut_noise, stack_kern, padding_size))
            amplification += projection
            delta = self.update_delta(delta, data, momentum, self.beta * self.alpha, projection)
        return delta.detach()


class RefGRA(RefAttack):              # gradient/gra.py:33-153
    def __init__(self, model, beta=3.5, num_neighbor=20, **kw):
        super().__init__(model, **kw)
        self.radius, self.num_neighbor = beta * self.epsilon, num_neighbor

    def forward(self, data, label, **kw):
        data, label = self._prep(data, label)
        delta = self.init_delta(data)
        eta = 0.94
        M = torch.full_like(delta, 1 / eta)
        momentum = 0
        for _ in range(self.epoch):
            grad = self.get_grad(self.get_loss(self.get_logits(self.transform(data + delta, momentum=momentum)), label), delta)
            sam = 0
            for _k in range(self.num_neighbor):                                                           # gra.py:42-58
                noise = torch.zeros_like(delta).uniform_(-self.radius, self.radius).to(self.device)
                sam += self.get_grad(self.get_loss(self.get_logits(self.transform(data + delta + noise, momentum=momentum)), label), delta)
            sam = sam / self.num_neighbor
            a, b = grad.view(grad.size(0), -1), sam.view(sam.size(0), -1)                                 # gra.py:60-72
            s = (torch.sum(a * b, dim=1) / (torch.sqrt(torch.sum(a ** 2, dim=1)) * torch.sqrt(torch.sum(b ** 2, dim=1)))).view(-1, 1, 1, 1)
            cur = s * grad + (1 - s) * sam
            last = momentum
            momentum = self.get_momentum(cur, momentum)
            last_t = torch.full(momentum.shape, last).to(momentum.device) if isinstance(last, int) else last   # gra.py:79-85
            eq = (last_t.sign() == momentum.sign()).float()                                              # gra.py:87-91
            M = M * (eq + (torch.ones_like(delta) - eq) * eta)
            delta = self.update_delta(delta, data, momentum, M * self.alpha)
        return delta.detach()


class RefAdaEA(RefAttack):            # ensemble/adaea.py:10-150 (model: RefEnsemble)
    def __init__(self, model, beta=10, threshold=-0.3, random_start=True, **kw):
        super().__init__(model, random_start=random_start, **kw)
        self.beta, self.threshold, self.K = beta, threshold, model.num_models

    def _one_step(self, x0, xa, g):                                                                     # adaea.py:138-148
        d = torch.clamp(xa.detach() + g.sign() * self.alpha - x0.detach(), -self.epsilon, self.epsilon)
        return torch.clamp(x0.detach() + d, max=1.0, min=0.0)

    def drf_map(self, grads, shape):                                                                    # adaea.py:115-136
        K, (B, _, H, W) = self.K, shape
        pair = torch.zeros(K, K, B, H, W, dtype=torch.float, device=self.device)
        rows = torch.zeros(K, B, H, W, dtype=torch.float, device=self.device)
        cos = nn.CosineSimilarity(dim=1, eps=1e-8)
        for i in range(K):
            for j in range(i + 1, K):
                pair[i][j] = cos(F.normalize(grads[i], dim=1), F.normalize(grads[j], dim=1))
            if i < K - 1:                       # the reference tests the inner loop's leaked j (= K-1)
                rows[i] = (pair[i, :].sum(dim=0) + pair[:, i].sum(dim=0)) / (K - 1)
        return rows.mean(dim=0).view(B, 1, H, W)

    def forward(self, data, label, **kw):
        data, label = data.clone().detach().to(self.device), label.clone().detach().to(self.device)
        ce = nn.CrossEntropyLoss()
        K, members = self.K, self.model.models
        momentum = 0.
        delta = torch.zeros_like(data).to(self.device) + 0.001 * torch.randn(data.shape, device=self.device)
        delta.requires_grad = True
        for _ in range(self.epoch):
            outs = [members[k](delta + data) for k in range(K)]
            grads = [torch.autograd.grad(ce(outs[k], label), delta, retain_graph=True, create_graph=False)[0] for k in range(K)]
            adv = [self._one_step(data, data + delta, grads[k]) for k in range(K)]                       # agm, adaea.py:87-113
            own = [ce(members[k](adv[k]), label) for k in range(K)]
            w = torch.zeros(size=(K,), device=self.device)
            for j in range(K):
                for i in range(K):
                    if i != j:
                        w[j] += ce(members[i](adv[j]), label) / own[i] * self.beta
            w = torch.softmax(w, dim=0)
            mp = self.drf_map(grads, data.shape)
            mp[mp >= self.threshold] = 1.
            mp[mp < self.threshold] = 0.
            out = (torch.stack(outs, dim=0) * w.view(K, 1, 1)).sum(dim=0)
            grad = torch.autograd.grad(ce(out, label).sum(dim=0), delta)[0] * mp
            momentum = self.get_momentum(grad, momentum)
            delta = self.update_delta(delta, data, momentum, self.alpha)
        return delta.detach()


class RefSSM(RefAttack):              # input_transformation/ssm.py:8-200 (device-agnostic: the reference hard-codes .cuda())
    def __init__(self, model, num_spectrum=20, rho=0.5, **kw):
        super().__init__(model, **kw)
        self.num_spectrum, self.rho = num_spectrum, rho

    @staticmethod
    def dct(x):                                                                                      # ssm.py:101-133, norm=None
        shape, N = x.shape, x.shape[-1]
        x = x.contiguous().view(-1, N)
        Vc = torch.fft.fft(torch.cat([x[:, ::2], x[:, 1::2].flip([1])], dim=1))
        k = -torch.arange(N, dtype=x.dtype, device=x.device)[None, :] * np.pi / (2 * N)
        return 2 * (Vc.real * torch.cos(k) - Vc.imag * torch.sin(k)).view(*shape)

    @staticmethod
    def idct(X):                                                                                     # ssm.py:135-172, norm=None
        shape, N = X.shape, X.shape[-1]
        Xv = X.contiguous().view(-1, N) / 2
        k = torch.arange(N, dtype=X.dtype, device=X.device)[None, :] * np.pi / (2 * N)
        Wr, Wi = torch.cos(k), torch.sin(k)
        Vti = torch.cat([Xv[:, :1] * 0, -Xv.flip([1])[:, :-1]], dim=1)
        v = torch.fft.ifft(torch.complex(real=Xv * Wr - Vti * Wi, imag=Xv * Wi + Vti * Wr))
        x = v.new_zeros(v.shape)
        x[:, ::2] += v[:, :N - (N // 2)]
        x[:, 1::2] += v.flip([1])[:, :N // 2]
        return x.view(*shape).real

    def dct_2d(self, x):
        return self.dct(self.dct(x).transpose(-1, -2)).transpose(-1, -2)

    def idct_2d(self, X):
        return self.idct(self.idct(X).transpose(-1, -2)).transpose(-1, -2)

    def transform(self, x, **kw):                                                                    # ssm.py:41-55
        gauss = (torch.randn(x.size()[0], 3, 224, 224) * self.epsilon).to(x.device)
        x_dct = self.dct_2d(x + gauss)
        mask = torch.rand_like(x) * 2 * self.rho + 1 - self.rho
        return self.idct_2d(x_dct * mask)

    def forward(self, data, label, **kw):
        data, label = self._prep(data, label)
        delta = self.init_delta(data)
        momentum = 0
        for _ in range(self.epoch):
            grads = 0
            for _k in range(self.num_spectrum):
                x_idct = self.transform(data + delta)
                grads += self.get_grad(self.get_loss(self.get_logits(x_idct), label), x_idct)
            grads /= self.num_spectrum
            momentum = self.get_momentum(grads, momentum)
            delta = self.update_delta(delta, data, momentum, self.alpha)
        return delta.detach()


REF_ZOO = {
    "fgsm": ref_fgsm, "ifgsm": ref_ifgsm, "mifgsm": ref_mifgsm, "nifgsm": RefNIFGSM, "dim": RefDIM,
    "tim": RefTIM, "sim": RefSIM, "admix": RefAdmix, "ditimi": RefDITIMI, "vmifgsm": RefVMIFGSM,
    "vnifgsm": RefVNIFGSM, "emifgsm": RefEMIFGSM, "ens": ref_mifgsm, "pifgsm": RefPIFGSM, "siditimi": RefSIDITIMI,
    "gra": RefGRA, "adaea": RefAdaEA, "ssm": RefSSM,
}


def save_images_u8(data, delta):
    """utils.py:64 — (adversaries.permute(0,2,3,1).cpu().numpy() * 255).astype(np.uint8)."""
    return ((data + delta).detach().permute((0, 2, 3, 1)).cpu().numpy() * 255).astype(np.uint8)
