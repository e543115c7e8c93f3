"""The ``Attack`` plugin base: the iterative loop and its eight overridable hooks.

Same class name, constructor, hook names/signatures, return types and error behaviour as the reference's
``transferattack/attack.py`` (lines cited per method), so existing attack subclasses run unchanged on top
of it. What is different is underneath: every per-iteration op around the surrogate's forward/backward is
a hand-written sm_100a kernel reached through ``ops`` (C-ABI ``libta_b200.so``):

  reference eager op chain                       here
  ---------------------------------------------  ------------------------------------------------------------
  data + delta                    (attack.py:88)  ``ta_stage_add`` inside an identity-backward autograd node,
                                                  or the ``xadv`` output of the previous fused update
  get_momentum: 5 ATen kernels   (attack.py:128)  ``ta_momentum`` (+ ``ta_abs_mean_per_sample`` / torch's mean)
  update_delta: 8 ATen kernels   (attack.py:147)  ``ta_update_linf`` / ``ta_update_l2``
  momentum+update+next data+delta                 ONE ``ta_fused_update_linf`` launch when neither hook is
                                                  overridden (the base loop owns delta/momentum, so in-place)
  init_delta clamp               (attack.py:141)  ``ta_clamp_box`` / ``ta_init_l2_scale``

The surrogate forward and ``torch.autograd.grad`` backward stay PyTorch. There is no CPU path: the hooks
raise on CPU tensors (``ops``), exactly like a missing ``libta_b200.so`` does.

``mean_mode`` selects how ``mean(|grad|)`` per sample is formed:
  'torch' (default) — the bits of the reference's own ``grad.abs().mean(dim=(1,2,3))``: formed INSIDE the fused kernel by
                      replaying the launch policy and fp32 summation tree of torch's CUDA mean kernel (``TA_MEAN_TORCH``,
                      csrc/aten_mean.cuh), so the whole tail is one launch and momentum / perturbation stay bit-identical
                      to the reference given the same gradient. The replay is self-checked against torch once per
                      (device, shape) (``ops.aten_mean_replay_ok``); where it does not apply, torch's own op supplies the scale;
  'aten'            — always torch's own ATen reduction for the scale (two more launches, +12 B/elem);
  'exact'           — reduced inside the kernels in fp64 (``TA_MEAN_EXACT``): correctly rounded mean, may differ from torch's
                      fp32 tree sum in the last bit.
"""
import warnings

import os

import torch
import torch.nn as nn

from . import _lib, ops
from .utils import *  # noqa: F401,F403  (plugins expect the reference's star-exports through this module too)
from .utils import EnsembleModel, PreprocessingModel, clamp, img_max, img_min, models, timm, wrap_model

_ZERO_TYPES = (int, float)


def _is_zero_scalar(v):
    return isinstance(v, _ZERO_TYPES) and not isinstance(v, bool) and v == 0


def _fold_bn(module):
    """In place on a COPY of the surrogate: every eval-mode BatchNorm2d that directly follows a Conv2d — `convK`/`bnK` attribute
    pairs of one parent (torchvision ResNet / Inception blocks) or neighbours inside an nn.Sequential (downsample, VGG-BN,
    MobileNet ConvNormActivation) — is folded into the convolution's weights and bias (torch.nn.utils.fusion) and replaced by
    Identity. Exact algebra, different rounding: fast mode only."""
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    n = 0
    for parent in list(module.modules()):
        kids = dict(parent.named_children())
        if isinstance(parent, nn.Sequential):
            names = list(kids)
            for a, b in zip(names, names[1:]):
                ca, cb = getattr(parent, a), getattr(parent, b)
                if isinstance(ca, nn.Conv2d) and isinstance(cb, nn.BatchNorm2d) and cb.track_running_stats:
                    setattr(parent, a, fuse_conv_bn_eval(ca, cb)); setattr(parent, b, nn.Identity()); n += 1
        for name, child in kids.items():
            if name.startswith("conv") and isinstance(child, nn.Conv2d):
                bn_name = "bn" + name[4:]
                bn = kids.get(bn_name)
                if isinstance(bn, nn.BatchNorm2d) and bn.track_running_stats and isinstance(getattr(parent, name), nn.Conv2d):
                    setattr(parent, name, fuse_conv_bn_eval(child, bn)); setattr(parent, bn_name, nn.Identity()); n += 1
    return n


class _FastMember(nn.Module):
    """fast mode only: Preprocessing (fp32 kernels) → a private copy of the network (BatchNorm folded into the convolutions and /
    or bf16 channels_last, per `mode`) → fp32 logits"""

    def __init__(self, wrapped, mode):
        super().__init__()
        import copy
        if isinstance(wrapped, nn.Sequential) and len(wrapped) == 2 and isinstance(wrapped[0], PreprocessingModel):
            self.pre, net = wrapped[0], wrapped[1]
        else:
            self.pre, net = None, wrapped
        net = copy.deepcopy(net).eval()
        self.folded = _fold_bn(net) if "bnfold" in mode else 0
        self.bf16 = "bf16" in mode
        if self.bf16:
            net = net.to(dtype=torch.bfloat16).to(memory_format=torch.channels_last)
        self.net = net
        for p_ in self.net.parameters():
            p_.requires_grad_(False)

    def forward(self, x):
        h = x if self.pre is None else self.pre(x)
        if self.bf16:
            h = h.to(torch.bfloat16)
            if h.dim() == 4:
                h = h.contiguous(memory_format=torch.channels_last)
        return self.net(h).float()


def _fast_twin(model, mode):
    if isinstance(model, EnsembleModel):
        return EnsembleModel([_FastMember(m, mode) for m in model.models], mode=model.mode)
    return _FastMember(model, mode)


class Attack(object):
    """Base class of every attack plugin (reference attack.py:8-169)."""

    #: 'torch' | 'aten' | 'exact' — see module docstring. Env override: TA_B200_MEAN.
    mean_mode = os.environ.get("TA_B200_MEAN", "torch")
    #: use the single-launch fused tail in the base loop when the hooks are not overridden
    fuse_update = os.environ.get("TA_B200_FUSE", "1") != "0"
    #: capture one iteration of the fused loop (staging → surrogate fwd/bwd → fused update) in a CUDA graph and replay it
    #: `epoch` times per batch: removes the ~550 host launches per iteration (measured on B200, profiles/graph_vs_eager_r1.json:
    #: +10 % at ResNet-50 B=64, 2.0x at B=8). Same kernels, same order → same bits (tests/test_e2e_gpu.py). On by default;
    #: a surrogate that cannot be captured (host syncs, data-dependent control flow) makes the loop fall back to launching
    #: the very same kernels eagerly. Env TA_B200_GRAPH=0 disables.
    use_cuda_graph = os.environ.get("TA_B200_GRAPH", "1") == "1"
    #: Declared (in the class BODY) by every class that defines loop hooks which are safe to capture once and replay:
    #: no host-side random draws or data-dependent Python control flow per call. A hook defined by a class that does not
    #: declare it — e.g. the reference's own dim.py on this base, whose transform flips a host coin per call — keeps the
    #: loop eager, so capture can never freeze such a decision into a graph.
    graph_safe = True
    #: captured graphs kept per attacker (one per batch shape); the oldest is dropped beyond this
    max_cached_graphs = 4
    #: SURVEY §8 f1: when the surrogate is ``Sequential(PreprocessingModel, net)`` (what ``wrap_model`` builds), its Resize is a
    #: no-op at the input size and neither ``transform`` nor ``get_logits`` is overridden, the fused tail writes the NORMALISED
    #: next input ((data + delta') - mean) / std itself (``ta_fused_update_linf_nf``) and ``net`` is entered directly: the
    #: Normalize forward kernel disappears from every iteration. Its adjoint g / std stays ONE launch at the end of the backward
    #: pass (``ta_normalize_bwd``; by default the variant that also finishes mean|g|, see ``colsum_adjoint``) or moves into the
    #: fused kernel too (``fold_adjoint``; default in 'exact' mode with the base ``get_grad``). Same arithmetic in the same
    #: order → same bits. Env TA_B200_FOLD=0 disables.
    fold_normalize = os.environ.get("TA_B200_FOLD", "1") == "1"
    #: with an in-kernel mean and the base get_grad, Normalize's ADJOINT (g / std) can be applied inside the tail kernels too
    #: instead of as a `ta_normalize_bwd` launch at the end of the backward pass. Same bits either way. Measured on B200 at B = 64
    #:: folded = 2 launches, 64 us of tail; not folded = adjoint kernel (inside autograd.grad) + 45-55 us of tail —
    #: the IEEE division has to be done in the mean kernel AND in the streaming kernel when folded. Default: not folded for
    #: mean_mode 'torch' (the adjoint kernel finishes the mean, see below), folded for 'exact' (one cluster launch).
    fold_adjoint = {"1": True, "0": False}.get(os.environ.get("TA_B200_FOLD_ADJOINT", ""), None)
    #: with mean_mode 'torch', the folded Normalize and the base get_grad: the Normalize-adjoint kernel at the end of the backward
    #: pass also forms the per-column sums of |g| of torch's mean reduction and its last CTA per sample finishes mean|g| from them
    #: (``ta_normalize_bwd_colsum``): the tail is the streaming kernel alone and the gradient is not read a second time for the mean.
    #: Same bits (self-checked per device and shape against torch's ops); False keeps the separate mean kernel.
    colsum_adjoint = os.environ.get("TA_B200_COLSUM_ADJOINT", "1") == "1"
    #: OPT-IN, NOT THE PARITY PATH. The surrogate's forward/backward runs on a private copy of
    #: the model: 'bnfold' = every eval-mode BatchNorm folded into its convolution (the launch list shows BN inference + BN backward
    #: at 30 % of an iteration), 'bf16' = bf16 / channels_last, 'bnfold+bf16' = both; everything around it — staging, mean|g|,
    #: momentum, update, clipping — stays the fp32 kernels. The perturbation is a valid one (eps-ball, [0,1] box) of the same attack
    #: but NOT bit-comparable with the reference: acceptance is attack strength (tests/test_e2e_gpu.py, bench.py `fast_mode`),
    #: never 1e-5 / uint8 identity. Off by default; env TA_B200_FAST enables. VMI/VNI additionally batch their neighbour
    #: evaluations in this mode (`fast_neighbor_images` images per forward).
    fast_mode = os.environ.get("TA_B200_FAST", "")
    fast_neighbor_images = 512

    def __init__(self, attack, model_name, epsilon, targeted, random_start, norm, loss, device=None):
        """attack.py:12-38 — same arguments, same attributes, same ``Unsupported norm`` exception."""
        if norm not in ['l2', 'linfty']:
            raise Exception("Unsupported norm {}".format(norm))
        self.attack = attack
        self.model = self.load_model(model_name)
        self.epsilon = epsilon
        self.targeted = targeted
        self.random_start = random_start
        self.norm = norm
        if isinstance(self.model, EnsembleModel):
            self.device = self.model.device
        else:
            self.device = next(self.model.parameters()).device if device is None else device
        self.loss = self.loss_function(loss)

    # ------------------------------------------------------------------------------------------------
    def load_model(self, model_name):
        """attack.py:40-65 — torchvision first, then timm; ``.eval().cuda()``; list → EnsembleModel.
        Subclasses with customised surrogates override this (documented override point)."""
        def load_single_model(name):
            if name in models.__dict__.keys():
                print('=> Loading model {} from torchvision.models'.format(name))
                model = models.__dict__[name](weights="DEFAULT")
            elif name in timm.list_models():
                print('=> Loading model {} from timm.models'.format(name))
                model = timm.create_model(name, pretrained=True)
            else:
                raise ValueError('Model {} not supported'.format(name))
            return wrap_model(model.eval().cuda())

        if isinstance(model_name, list):
            return EnsembleModel([load_single_model(name) for name in model_name])
        return load_single_model(model_name)

    # ------------------------------------------------------------------------------------------------
    def _to_device(self, t):
        """attack.py:79-80 clones then moves. The kernels never write into ``data``/``label``, so a tensor that
        is already on the device is used as is, and host tensors go up with one (async if pinned) copy."""
        t = t.detach()
        if t.device == torch.device(self.device) or (t.is_cuda and torch.device(self.device).index is None):
            return t
        return t.to(self.device, non_blocking=t.is_pinned() if not t.is_cuda else False)

    #: hooks whose OWNER class must itself declare graph_safe = True for the loop to be captured. ``load_model`` is one of
    #: them: it is the reference's documented override point for customised surrogates (sapr, ghost, sgm, ... override only
    #: it), and a surrogate with host-side randomness or Python control flow in its forward must never be frozen into a graph.
    _GRAPH_HOOKS = ("forward", "transform", "get_logits", "get_loss", "get_grad", "get_momentum", "update_delta", "init_delta",
                    "load_model")

    def _graph_ok(self):
        """every loop hook in effect (and the surrogate's loader) is defined by a class that itself declares graph_safe = True,
        and the surrogate carries no forward / backward module hooks (registered by code that did not opt in)"""
        for hook in self._GRAPH_HOOKS:
            owner = next(c for c in type(self).__mro__ if hook in c.__dict__)
            if not owner.__dict__.get("graph_safe", False):
                return False
        if not getattr(self, "graph_safe_module_hooks", False) and isinstance(self.model, nn.Module):
            for mod in self.model.modules():
                if (mod._forward_hooks or mod._forward_pre_hooks or mod._backward_hooks
                        or getattr(mod, "_backward_pre_hooks", None)):
                    return False
        return True

    def _mean_kernel_mode(self, like):
        """the in-kernel mean mode for gradients shaped like `like`, or None = take the scale from torch's own op"""
        if self.mean_mode == 'exact':
            return _lib.TA_MEAN_EXACT
        if self.mean_mode == 'torch' and ops.aten_mean_replay_ok(like):
            return _lib.TA_MEAN_TORCH
        return None

    def _surrogate(self):
        """the module get_logits runs: `self.model`, or in fast mode its bf16 / channels_last twin (built once per model)"""
        if not self.fast_mode:
            return self.model
        if self.fast_mode not in ('bnfold', 'bf16', 'bnfold+bf16'):
            raise ValueError("unknown fast_mode {!r} ('bnfold', 'bf16' or 'bnfold+bf16')".format(self.fast_mode))
        cached = self.__dict__.get("_fast_twin")
        if cached is None or cached[0] is not self.model or cached[2] != self.fast_mode:
            cached = (self.model, _fast_twin(self.model, self.fast_mode), self.fast_mode)
            self.__dict__["_fast_twin"] = cached
        return cached[1]

    def _fusable(self):
        cls = type(self)
        return (self.fuse_update and self.norm == 'linfty'
                and cls.get_momentum is Attack.get_momentum and cls.update_delta is Attack.update_delta
                and cls.init_delta is Attack.init_delta
                and isinstance(self.alpha, (int, float)) and isinstance(self.decay, (int, float)))

    def _fold_plan(self, data, kmode=None):
        """(pre, net, mean, std, defer, colsum) when Normalize can be folded into the fused tail for this batch, else None.
        defer: Normalize's adjoint is applied inside the tail kernels; colsum: the adjoint kernel leaves the column sums of |g|.
        `kmode`: the in-kernel mean mode (``_mean_kernel_mode``); with one, Normalize's adjoint moves into the kernel too."""
        cls = type(self)
        if not self.fold_normalize or self.fast_mode or cls.get_logits is not Attack.get_logits or cls.transform is not Attack.transform:
            return None
        m = self.model
        if not (isinstance(m, nn.Sequential) and len(m) == 2 and isinstance(m[0], PreprocessingModel)) or data.dim() != 4:
            return None
        pre = m[0]
        B, C, H, W = data.shape
        size = pre.resize.size
        size = size[0] if isinstance(size, (list, tuple)) and len(size) == 1 else size
        if not isinstance(size, int) or min(H, W) != size:           # Resize(int) keeps the tensor only when the short side matches
            return None
        if pre.mean.numel() != C or C > 4 or (H * W) % 4 != 0 or data.data_ptr() % 16 != 0:
            return None
        # Normalize's adjoint inside the kernel needs the staged (cluster) form: the sample must fit 8 CTAs' shared memory
        fa = self.fold_adjoint if self.fold_adjoint is not None else (kmode == _lib.TA_MEAN_EXACT)
        defer = fa and kmode is not None and cls.get_grad is Attack.get_grad and C * H * W <= 384 * 1024
        colsum = False
        if self.colsum_adjoint and not defer and kmode == _lib.TA_MEAN_TORCH and cls.get_grad is Attack.get_grad:
            pre._buffers_on(data.device)
            colsum = ops.colsum_adjoint_ok(data, pre.std)
        return pre, m[1], [float(v) for v in pre.mean.tolist()], [float(v) for v in pre.std.tolist()], defer, colsum

    @staticmethod
    def _first_normalized(pre, data, delta, out=None):
        """xn of the first iteration: the two separate kernels, once per batch"""
        be = ops.backend()
        with torch.no_grad():
            pre._buffers_on(data.device)
            xn = be.normalize(be.stage_add(data, delta.detach()), pre.mean, pre.std, True)
            if out is not None:
                out.copy_(xn)
                return out
        return xn

    def forward(self, data, label, **kwargs):
        """The general attack procedure (attack.py:67-102).

        data (N, C, H, W); label (N,) or (2, N) = [ground truth, target] when targeted. Returns delta.detach().
        """
        if self.targeted:
            assert len(label) == 2
            label = label[1]
        data = self._to_device(data).contiguous()
        label = self._to_device(label)

        delta = self.init_delta(data)
        if self._fusable():
            if (self.use_cuda_graph and self._graph_ok() and data.is_cuda and ops._test_backend is None
                    and getattr(self, "_kernel_events", None) is None and not self.__dict__.get("_graph_failed", False)):
                self._mean_kernel_mode(data)     # the one-time self-check synchronises: never inside the capture
                try:
                    return self._loop_graph(data, label, delta)
                except RuntimeError as e:
                    # only a REFUSED CAPTURE (the surrogate synchronises, allocates through an uncapturable path, ...) turns
                    # the loop eager; out-of-memory, kernel failures and bugs in hooks are raised as they are
                    msg = str(e)
                    if isinstance(e, torch.OutOfMemoryError) or "libta_b200" in msg or not any(
                            k in msg.lower() for k in ("captur", "cudagraph", "cuda graph", "graph")):
                        raise
                    self._graph_failed = True
                    self._graph_error = msg
                    warnings.warn("transferattack_b200: CUDA-graph capture of the attack iteration was refused (%s); "
                                  "launching the same kernels eagerly from now on" % msg.splitlines()[0][:200])
                    torch.cuda.synchronize(data.device)
            return self._loop_fused(data, label, delta)

        momentum = 0
        for _ in range(self.epoch):
            logits = self.get_logits(self.transform(ops.stage_add(data, delta), momentum=momentum))
            loss = self.get_loss(logits, label)
            grad = self.get_grad(loss, delta)
            momentum = self.get_momentum(grad, momentum)
            delta = self.update_delta(delta, data, momentum, self.alpha)
        return delta.detach()

    def _tail(self, be, grad, momentum, m_out, delta, delta_out, data, xadv, scale_out, kmode, fold, addend=None, gbar_out=None,
              col_sums=None):
        """get_momentum + update_delta + the next model input as ONE ``ta_fused_tail`` launch. `kmode` None (or a shape the
        in-kernel reduction does not serve): the scale comes from torch's own ``abs().mean`` op and the streaming form runs."""
        norm = {}
        if fold is not None:
            norm = dict(mean=fold[2], std=fold[3], emit_normalized=True, grad_wrt_xn=fold[4])
        with torch.no_grad():
            if col_sums is not None:       # mean|g| was finished by the adjoint kernel (torch's bits) into scale_out: the streaming form
                if not be.fused_tail(grad, momentum, m_out, delta, delta_out, data, xadv, scale_out, scale_out, self.decay, self.alpha,
                                     self.epsilon, img_min, img_max, gbar_out=gbar_out, **norm):
                    raise RuntimeError("ta_fused_tail refused the streaming form: %s" % _lib.last_error())
                return
            if kmode is not None and be.fused_tail(grad, momentum, m_out, delta, delta_out, data, xadv, None, scale_out, self.decay,
                                                   self.alpha, self.epsilon, img_min, img_max, mean_mode=kmode, addend=addend,
                                                   gbar_out=gbar_out, **norm):
                return
            if fold is not None and fold[4]:
                raise RuntimeError("ta_fused_tail refused a folded shape _fold_plan accepted")
            g = grad if addend is None else be.add(grad, addend)
            if not be.fused_tail(g, momentum, m_out, delta, delta_out, data, xadv, self._torch_abs_mean(g), scale_out, self.decay,
                                 self.alpha, self.epsilon, img_min, img_max, gbar_out=gbar_out, **norm):
                raise RuntimeError("ta_fused_tail refused the streaming form: %s" % _lib.last_error())

    def _loop_fused(self, data, label, delta):
        """attack.py:86-100 with get_momentum + update_delta + the next `data + delta` in one launch per
        iteration. delta / momentum / x_adv live in buffers this loop owns and are updated in place."""
        be = ops.backend()
        m_buf = torch.empty_like(data)
        scale_out = torch.empty(data.shape[0], device=data.device, dtype=torch.float32)
        kmode = self._mean_kernel_mode(data)
        fold = self._fold_plan(data, kmode)
        col_sums = None
        if fold is not None:
            pre, net, mean, std, defer, colsum = fold
            xadv = self._first_normalized(pre, data, delta)          # holds the NORMALISED model input from here on
            if colsum:                  # (column sums, where the adjoint kernel leaves mean|g|, its ticket counters)
                col_sums = (torch.empty(data.shape[0] * be.colsum_size(data.shape[0], data[0].numel(), data.device), device=data.device,
                                        dtype=torch.float32), scale_out, torch.zeros(data.shape[0], device=data.device, dtype=torch.int32))
        else:
            xadv = torch.empty_like(data)
        momentum, pre_x = None, None
        for _ in range(self.epoch):
            if fold is not None:
                logits = net(ops.stage_normalized(delta, xadv, pre.std, defer, col_sums))
            else:
                x = ops.stage_add(data, delta, precomputed=pre_x)
                logits = self.get_logits(self.transform(x, momentum=0 if momentum is None else momentum))
            loss = self.get_loss(logits, label)
            grad = self.get_grad(loss, delta)
            ev = getattr(self, "_kernel_events", None)      # bench.py: CUDA events around the WHOLE tail (everything between
            if ev is not None:                               # autograd.grad and the next forward), on its stream
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            self._tail(be, grad, momentum, m_buf, delta, delta, data, xadv, scale_out, kmode, fold, col_sums=col_sums)
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
            momentum, pre_x = m_buf, xadv
        return delta.detach()

    # ---- CUDA-graph replay of the fused loop -------------------------------------------------------------
    def _graph_iteration(self, st):
        """One iteration on the static buffers `st` (this body is what gets captured)."""
        fold = st.get("fold")
        if fold is not None:
            pre, net, mean, std, defer, colsum = fold
            logits = net(ops.stage_normalized(st["delta"], st["xadv"], pre.std, defer, st.get("col_sums")))
        else:
            x = ops.stage_add(st["data"], st["delta"], precomputed=st["xadv"])
            logits = self.get_logits(self.transform(x, momentum=st["m"]))
        loss = self.get_loss(logits, st["label"])
        grad = self.get_grad(loss, st["delta"])
        self._tail(ops.backend(), grad, st["m"], st["m"], st["delta"], st["delta"], st["data"], st["xadv"], st["scale_out"],
                   st["kmode"], fold, col_sums=st.get("col_sums"))
        step = getattr(self, "_graph_step", None)       # plugins with per-iteration state on the device (DIM's draw index)
        if step is not None:
            step()

    def _graph_reset(self, st, data, label, delta0):
        with torch.no_grad():
            st["data"].copy_(data)
            st["label"].copy_(label)
            st["delta"].copy_(delta0)
            st["m"].zero_()          # momentum * decay with momentum = +0 is the reference's first-iteration `0 * decay`
            if st.get("fold") is not None:
                self._first_normalized(st["fold"][0], st["data"], st["delta"], out=st["xadv"])
            else:
                ops.backend().stage_add(st["data"], st["delta"], out=st["xadv"])
            rewind = getattr(self, "_graph_rewind", None)
            if rewind is not None:
                rewind()

    def _graph_for(self, data, label, delta0):
        kmode = self._mean_kernel_mode(data)
        fold = self._fold_plan(data, kmode)
        key = (tuple(data.shape), str(data.device), tuple(label.shape), self.mean_mode, kmode, float(self.alpha), float(self.decay),
               float(self.epsilon), bool(self.targeted), id(self.model), fold is not None, bool(fold[4]) if fold else False,
               bool(fold[5]) if fold else False, self.fast_mode)
        cache = self.__dict__.setdefault("_graphs", {})
        st = cache.get(key)
        if st is not None:
            return st
        while len(cache) >= self.max_cached_graphs:      # each graph pins its static buffers + the surrogate's activation pool
            cache.pop(next(iter(cache)))
        st = {"data": torch.empty_like(data), "label": torch.empty_like(label),
              "delta": torch.zeros_like(data).requires_grad_(True), "m": torch.zeros_like(data),
              "xadv": torch.empty_like(data), "scale_out": torch.empty(data.shape[0], device=data.device, dtype=torch.float32),
              "fold": fold, "kmode": kmode}
        if fold is not None and fold[5]:
            be = ops.backend()
            st["col_sums"] = (torch.empty(data.shape[0] * be.colsum_size(data.shape[0], data[0].numel(), data.device), device=data.device,
                                          dtype=torch.float32), st["scale_out"], torch.zeros(data.shape[0], device=data.device, dtype=torch.int32))
        self._graph_reset(st, data, label, delta0)
        cur = torch.cuda.current_stream(data.device)
        side = torch.cuda.Stream(device=data.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up off the capture: cuDNN heuristics, workspaces, our smem attributes
            for _ in range(3):
                self._graph_iteration(st)
        cur.wait_stream(side)
        torch.cuda.synchronize(data.device)
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count() if ops._test_backend is None else 0
        with torch.cuda.graph(graph):
            self._graph_iteration(st)
        st["graph"] = graph
        st["kernels_per_replay"] = (_lib.launch_count() - n0) if ops._test_backend is None else 0   # OUR kernels inside one replay
        cache[key] = st
        return st

    def _loop_graph(self, data, label, delta0):
        begin, end = getattr(self, "_graph_begin", None), getattr(self, "_graph_end", None)
        if begin is not None:                       # e.g. DIM: draw all `epoch` transforms now, in the reference's order
            begin(data)
        ok = False
        try:
            st = self._graph_for(data, label, delta0.detach())
            self._graph_reset(st, data, label, delta0.detach())
            for _ in range(self.epoch):
                st["graph"].replay()
            out = st["delta"].detach().clone()
            ok = True
            return out
        finally:
            if end is not None:
                end(ok)

    # ------------------------------------------------------------------------------------------------
    def get_logits(self, x, **kwargs):
        """attack.py:104-108 (fast mode: the bf16 twin, see ``fast_mode``)"""
        return self._surrogate()(x)

    def get_loss(self, logits, label):
        """attack.py:110-115"""
        return -self.loss(logits, label) if self.targeted else self.loss(logits, label)

    def get_grad(self, loss, delta, **kwargs):
        """attack.py:118-122 — the surrogate's backward (torch autograd; the staging kernels' adjoints are
        autograd nodes inside that graph)."""
        return torch.autograd.grad(loss, delta, retain_graph=False, create_graph=False)[0]

    @staticmethod
    def _torch_abs_mean(grad):
        return grad.abs().mean(dim=(1, 2, 3))

    def _abs_mean(self, grad):
        """mean|grad| per sample for the public hooks: one ``ta_abs_mean_per_sample`` launch in the attack's mean mode, or
        torch's own op where the in-kernel replay does not apply"""
        kmode = self._mean_kernel_mode(grad) if grad.dim() >= 2 else None
        if kmode is not None:
            out = ops.backend().abs_mean(grad, kmode)
            if out is not None:
                return out
        return self._torch_abs_mean(grad)

    def get_momentum(self, grad, momentum, **kwargs):
        """attack.py:124-128: momentum * decay + grad / mean(|grad|) per sample. ``momentum`` may be the
        Python 0 of the first iteration; unknown kwargs (e.g. ``decay=``) are ignored like in the reference."""
        m = None if _is_zero_scalar(momentum) else momentum
        if m is not None and not torch.is_tensor(m):
            raise TypeError("momentum must be a tensor or 0, got {}".format(type(momentum)))
        return ops.backend().momentum(grad, m, self._abs_mean(grad), self.decay)

    def init_delta(self, data, **kwargs):
        """attack.py:130-143. Random draws come from torch's device generator (same stream of numbers as the
        reference); the projection runs in ``ta_clamp_box`` / ``ta_init_l2_scale``."""
        delta = torch.zeros_like(data).to(self.device)
        if self.random_start:
            be = ops.backend()
            if self.norm == 'linfty':
                delta.uniform_(-self.epsilon, self.epsilon)
                delta = be.clamp_box(delta, data, img_min, img_max)
            else:
                delta.normal_(-self.epsilon, self.epsilon)
                r = torch.zeros_like(data).uniform_(0, 1).to(self.device)
                delta = be.init_l2_scale(delta, r, data, self.epsilon, img_min, img_max)
        delta.requires_grad = True
        return delta

    def update_delta(self, delta, data, grad, alpha, **kwargs):
        """attack.py:145-153. ``alpha`` may be a float (also negative) or a tensor broadcastable to delta.
        Returns a fresh leaf with requires_grad=True; the inputs are not modified."""
        be = ops.backend()
        if self.norm == 'linfty':
            if torch.is_tensor(alpha):
                if alpha.numel() == 1:
                    out = be.update_linf(delta, data, grad, float(alpha), self.epsilon, img_min, img_max)
                else:
                    a = alpha.to(device=delta.device, dtype=torch.float32).expand_as(delta)
                    out = be.update_linf(delta, data, grad, 0.0, self.epsilon, img_min, img_max, alpha_t=a)
            else:
                out = be.update_linf(delta, data, grad, alpha, self.epsilon, img_min, img_max)
        else:
            out = be.update_l2(delta, data, grad, alpha, self.epsilon, img_min, img_max)
        return out.detach().requires_grad_(True)

    def loss_function(self, loss):
        """attack.py:155-162"""
        if loss == 'crossentropy':
            return nn.CrossEntropyLoss()
        raise Exception("Unsupported loss {}".format(loss))

    def transform(self, data, **kwargs):
        """attack.py:164-165"""
        return data

    def __call__(self, *input, **kwargs):
        """attack.py:167-169"""
        self.model.eval()
        return self.forward(*input, **kwargs)
