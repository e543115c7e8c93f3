"""Shared helpers of the attack plugins — same names and star-export surface as the reference's
``transferattack/utils.py`` (plugins do ``from ..utils import *`` and rely on the re-exported
``torch, nn, models, transforms, np, pd, timm, os, Image``), re-implemented on the kernels.

Differences that are not visible to callers:
  * ``PreprocessingModel`` normalises with the ``ta_normalize_*`` kernels (bit-identical to torchvision's
    clone/sub_/div_ and without its per-forward ``(std == 0).any()`` host sync; reference utils.py:72-79);
  * ``clamp`` on CUDA fp32 tensors with per-element bounds of the box form is left to the callers' kernels;
    the function itself keeps the reference's tensor semantics (utils.py:68-69);
  * ``save_images`` quantises on the device with ``ta_quantize_u8`` before the single D2H copy (utils.py:63-66);
    ``AsyncImageWriter`` additionally takes the copy and the PNG encoding off the loop's critical path and
    ``PrefetchLoader`` uploads the next batch while the current one is attacked.
"""
import os
import types

import numpy as np
import torch
import torch.nn as nn
import torchvision.models as models
import torchvision.transforms as transforms
from PIL import Image

try:                                     # pandas is only needed by AdvDataset.load_labels
    import pandas as pd
except Exception:                        # pragma: no cover
    pd = None
try:
    import timm
except ModuleNotFoundError:              # timm is optional here; torchvision surrogates do not need it
    timm = types.ModuleType("timm")
    timm.list_models = lambda *a, **k: []

    def _no_timm(*a, **k):
        raise ModuleNotFoundError("timm is not installed; only torchvision surrogates are available")
    timm.create_model = _no_timm

from . import ops

img_height, img_width = 224, 224
img_max, img_min = 1., 0

cnn_model_paper = ['resnet50', 'vgg16', 'mobilenet_v2', 'inception_v3']
vit_model_paper = ['vit_base_patch16_224', 'pit_b_224', 'visformer_small', 'swin_tiny_patch4_window7_224']
cnn_model_pkg = ['vgg19', 'resnet18', 'resnet101', 'resnext50_32x4d', 'densenet121', 'mobilenet_v2']
vit_model_pkg = ['vit_base_patch16_224', 'pit_b_224', 'cait_s24_224', 'visformer_small',
                 'tnt_s_patch16_224', 'levit_256', 'convit_base', 'swin_tiny_patch4_window7_224']
tgr_vit_model_list = ['vit_base_patch16_224', 'pit_b_224', 'cait_s24_224', 'visformer_small',
                      'deit_base_distilled_patch16_224', 'tnt_s_patch16_224', 'levit_256', 'convit_base']
generation_target_classes = [24, 99, 245, 344, 471, 555, 661, 701, 802, 919]


def load_pretrained_model(cnn_model=[], vit_model=[]):
    """Generator of (name, model) pairs: torchvision weights for CNNs, timm for ViTs (utils.py:29-34)."""
    for name in cnn_model:
        yield name, models.__dict__[name](weights="DEFAULT")
    for name in vit_model:
        yield name, timm.create_model(name, pretrained=True)


class PreprocessingModel(nn.Module):
    """Resize (a no-op at the native size) then per-channel Normalize, the first stage of every wrapped
    surrogate. Normalisation runs in ``ta_normalize_fwd`` / ``ta_normalize_bwd``: (x - mean) / std with the
    reference's two roundings, adjoint g / std."""

    def __init__(self, resize, mean, std):
        super().__init__()
        self.resize = transforms.Resize(resize)
        self.register_buffer("mean", torch.tensor(list(mean), dtype=torch.float32), persistent=False)
        self.register_buffer("std", torch.tensor(list(std), dtype=torch.float32), persistent=False)
        if bool((self.std == 0).any()):          # checked once here instead of once per forward
            raise ValueError("std evaluated to zero, leading to division by zero.")

    def _buffers_on(self, device):
        if self.mean.device != device:
            self.mean = self.mean.to(device)
            self.std = self.std.to(device)

    def forward(self, x):
        x = self.resize(x)
        self._buffers_on(x.device)
        return ops.normalize(x, self.mean, self.std)


def wrap_model(model):
    """Prepend the training-time normalisation (utils.py:37-60): timm ``default_cfg`` mean/std; torchvision
    Inception → 0.5/0.5 at 299; everything else → ImageNet statistics at 224."""
    resize = 224
    if hasattr(model, 'default_cfg'):
        mean, std = model.default_cfg['mean'], model.default_cfg['std']
    elif 'Inc' in model.__class__.__name__:
        mean, std, resize = [0.5, 0.5, 0.5], [0.5, 0.5, 0.5], 299
    else:
        mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    return torch.nn.Sequential(PreprocessingModel(resize, mean, std), model)


def save_images(output_dir, adversaries, filenames, delta=None):
    """Write adversarial images as PNG. ``adversaries`` is the reference's ``images + perturbations`` tensor
    (utils.py:63-66: ``(x.permute(0,2,3,1).numpy() * 255).astype(uint8)``, i.e. truncation). On a CUDA tensor
    the multiply/truncate/NHWC transpose happens on the device (``ta_quantize_u8``) and only bytes cross PCIe;
    pass ``delta`` to fuse the ``images + perturbations`` sum as well."""
    if adversaries.is_cuda:
        zero = torch.zeros_like(adversaries) if delta is None else delta
        u8 = ops.backend().quantize_u8(adversaries, zero, to_nhwc=True).cpu().numpy()
    else:
        if delta is not None:
            adversaries = adversaries + delta
        u8 = (adversaries.detach().permute((0, 2, 3, 1)).cpu().numpy() * 255).astype(np.uint8)
    for i, filename in enumerate(filenames):
        Image.fromarray(u8[i]).save(os.path.join(output_dir, filename))


class AsyncImageWriter:
    """``save_images`` off the attack loop's critical path (reference utils.py:63-66 encodes every PNG serially
    between two batches). ``submit`` quantises on the device (``ta_quantize_u8``: add, *255, truncate, NHWC transpose — the
    reference's bytes), starts the device→host copy of the BYTES into a pinned buffer on a side stream and returns at once;
    a thread pool waits for the copy and encodes the PNGs in parallel (Pillow releases the GIL inside zlib). The next batch's
    attack overlaps both. ``flush`` (or leaving the ``with`` block) waits for everything and re-raises the first error."""

    def __init__(self, workers=8, max_pending=4):
        import concurrent.futures as cf
        self._pool = cf.ThreadPoolExecutor(max_workers=max(1, int(workers)))
        self._pending = []
        self._max_pending = max(1, int(max_pending))
        self._streams = {}

    @staticmethod
    def _encode(u8_row, path):
        Image.fromarray(u8_row).save(path)

    def _finish(self, event, host_u8, output_dir, filenames):
        if event is not None:
            event.synchronize()
        arr = host_u8.numpy()
        futs = [self._pool.submit(self._encode, arr[i], os.path.join(output_dir, fn)) for i, fn in enumerate(filenames)]
        for f in futs:
            f.result()

    def submit(self, output_dir, adversaries, filenames, delta=None):
        filenames = list(filenames)
        if adversaries.is_cuda:
            dev = adversaries.device
            zero = torch.zeros_like(adversaries) if delta is None else delta
            u8 = ops.backend().quantize_u8(adversaries, zero, to_nhwc=True)
            side = self._streams.get(dev.index)
            if side is None:
                side = self._streams[dev.index] = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            host = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=True)
            with torch.cuda.stream(side):
                host.copy_(u8, non_blocking=True)
                event = torch.cuda.Event()
                event.record(side)
            u8.record_stream(side)
        else:
            if delta is not None:
                adversaries = adversaries + delta
            host = torch.from_numpy((adversaries.detach().permute((0, 2, 3, 1)).cpu().numpy() * 255).astype(np.uint8))
            event = None
        import threading
        t = threading.Thread(target=self._run, args=(event, host, output_dir, filenames), daemon=True)
        t.err = None
        t.start()
        self._pending.append(t)
        while len(self._pending) > self._max_pending:       # bound the pinned memory in flight
            self._join(self._pending.pop(0))

    def _run(self, event, host, output_dir, filenames):
        import threading
        try:
            self._finish(event, host, output_dir, filenames)
        except BaseException as e:     # surfaced by flush()
            threading.current_thread().err = e

    @staticmethod
    def _join(t):
        t.join()
        if t.err is not None:
            raise t.err

    def flush(self):
        while self._pending:
            self._join(self._pending.pop(0))

    def close(self):
        self.flush()
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class PrefetchLoader:
    """Iterate a ``DataLoader`` one batch AHEAD on the device (reference main.py:41-52 hands the attack pageable
    host tensors and pays the upload inside every call): while the attack works on batch i, batch i+1 is already pinned
    (``pin_memory=True`` on the loader) and on its way up on a side stream. Yields ``(images_on_device, labels, filenames)``;
    labels stay where the loader put them (the attack moves the few bytes itself)."""

    def __init__(self, loader, device):
        self.loader = loader
        self.device = torch.device(device)
        self._side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def __len__(self):
        return len(self.loader)

    def _upload(self, batch):
        images = batch[0]
        if self._side is None:
            return (images.to(self.device),) + tuple(batch[1:]), None
        with torch.cuda.stream(self._side):
            dev = images.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._side)
        return (dev,) + tuple(batch[1:]), ev

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            try:
                nxt = self._upload(next(it))
            except StopIteration:
                nxt = None
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
                cur[0].record_stream(torch.cuda.current_stream(self.device))
            yield cur


def clamp(x, x_min, x_max):
    return torch.min(torch.max(x, x_min), x_max)


class EnsembleModel(torch.nn.Module):
    """Members evaluated one after another on one device; logits stacked and averaged (``mode='mean'``) or
    returned stacked (``mode='ind'``) — utils.py:82-105. Exposes ``models``, ``device``, ``num_models``,
    ``mode``, ``softmax``, ``type_name`` like the reference (SVRE/CWA/AdaEA index ``.models[k]``)."""

    def __init__(self, models, mode='mean'):
        super().__init__()
        self.device = next(models[0].parameters()).device
        for m in models:
            m.to(self.device)
        self.models = models
        self.softmax = torch.nn.Softmax(dim=1)
        self.type_name = 'ensemble'
        self.num_models = len(models)
        self.mode = mode

    def forward(self, x):
        outputs = torch.stack([m(x) for m in self.models], dim=0)
        if self.mode == 'mean':
            return torch.mean(outputs, dim=0)
        if self.mode == 'ind':
            return outputs
        raise NotImplementedError


class AdvDataset(torch.utils.data.Dataset):
    """``labels.csv`` + ``images/`` reader with the reference's item format ``(CHW float32 in [0,1], label,
    filename)`` (utils.py:108-153). The filename list is materialised once instead of per ``__getitem__``."""

    def __init__(self, input_dir=None, output_dir=None, targeted=False, target_class=None, eval=False):
        self.targeted = targeted
        self.target_class = target_class
        self.data_dir = input_dir
        self.f2l = self.load_labels(os.path.join(self.data_dir, 'labels.csv'))
        self._names = list(self.f2l.keys())
        if eval:
            self.data_dir = output_dir
            print('=> Eval mode: evaluating on {}'.format(self.data_dir))
        else:
            self.data_dir = os.path.join(self.data_dir, 'images')
            print('=> Train mode: training on {}'.format(self.data_dir))
            print('Save images to {}'.format(output_dir))

    def __len__(self):
        return len(self._names)

    def __getitem__(self, idx):
        filename = self._names[idx]
        assert isinstance(filename, str)
        image = Image.open(os.path.join(self.data_dir, filename))
        image = image.resize((img_height, img_width)).convert('RGB')
        image = torch.from_numpy(np.array(image).astype(np.float32) / 255).permute(2, 0, 1)
        return image, self.f2l[filename], filename

    def load_labels(self, file_name):
        dev = pd.read_csv(file_name)
        if self.targeted:
            second = (lambda i: self.target_class) if self.target_class else (lambda i: dev.iloc[i]['targeted_label'])
            return {dev.iloc[i]['filename']: [dev.iloc[i]['label'], second(i)] for i in range(len(dev))}
        return {dev.iloc[i]['filename']: dev.iloc[i]['label'] for i in range(len(dev))}
