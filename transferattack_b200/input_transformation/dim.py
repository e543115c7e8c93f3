"""DIM (Xie et al., CVPR 2019): with probability ``diversity_prob`` the input is bilinearly resized to a random size in
[S, int(S*resize_rate)), zero-padded at a random offset to int(S*resize_rate) and resized back to S.
Reference: transferattack/input_transformation/dim.py:35-68 (same constructor; the coin and the three integers are
drawn from torch's global CPU generator with the same calls in the same order, so a seeded run takes the same
decisions as the reference).

The three image-sized ops are ONE gather kernel (``ta_dim_fwd``: bulk-TMA-staged source rows, the intermediate
resize kept in shared memory) with a deterministic gather-form adjoint (``ta_dim_bwd``).

CUDA graph: a graph replays what ran at capture time, and DIM draws per call. So, for the base loop's graph path, ALL `epoch`
draws of an attack call are made up front — the same ``torch.rand`` / ``torch.randint`` calls in the same order as the
per-iteration draws of the eager loop (nothing else consumes the host generator in between) —, turned into per-iteration table
records on the host, uploaded once, and the captured kernels read record ``*it`` where ``it`` is a device counter advanced
inside the graph (``ta_dim_*_dyn``, ``ta_counter_add``). Same kernels, same tables → the same bits as the eager loop."""
from ..utils import *
from .. import ops
from ..gradient.mifgsm import MIFGSM


class DIM(MIFGSM):
    graph_safe = True       # the draws are made before the replays and reach the captured kernels through device memory (below)

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
        dyn = self.__dict__.get("_dim_dyn")
        if dyn is not None and dyn["active"]:          # inside the graph loop: the draw of iteration *it is already on the device
            return ops.dim_resize_pad_dyn(x, dyn["R"], dyn["packs"], dyn["n"], dyn["it"])
        pend = dyn.get("pending") if dyn is not None else None
        if pend:        # a graph loop drew this call's parameters and then had to fall back to eager launches: use them, in order
            d = pend.pop(0)
            params = None if d is None else (d[0], dyn["R"], d[1], d[2])
        else:
            params = self.draw(x.shape[-1])
        if params is None:
            return x
        return ops.dim_resize_pad(x, *params)

    # ---- hooks of Attack._loop_graph ------------------------------------------------------------------------------------
    def _graph_begin(self, data):
        """before warm-up / capture / replays of one attack call: draw every iteration's parameters (reference order), build the
        table records and upload them into this attacker's fixed device buffer"""
        be = ops.backend()
        S = data.shape[-1]
        R = int(S * self.resize_rate)
        draws = []
        for _ in range(self.epoch):
            p = self.draw(S)
            draws.append(None if p is None else (p[0], p[2], p[3]))
        dyn = self.__dict__.get("_dim_dyn")
        if dyn is None or dyn["cap"] < self.epoch or dyn["S"] != S or dyn["R"] != R or dyn["packs"].device != data.device:
            nb = int(be.lib.ta_dim_pack_bytes())
            cap = max(self.epoch, 16)
            dyn = {"packs": torch.empty((cap, nb), dtype=torch.uint8, device=data.device), "cap": cap, "S": S, "R": R,
                   "it": torch.zeros(1, dtype=torch.int32, device=data.device), "active": False}
            self.__dict__["_dim_dyn"] = dyn
            self.__dict__.pop("_graphs", None)          # captured graphs hold the old buffers' addresses
        host = be.dim_packs(draws, S, R)
        dyn["packs"][:self.epoch].copy_(host, non_blocking=True)
        dyn["host"] = host                               # keep the pinned staging alive until the copy has run
        dyn["n"] = self.epoch
        dyn["active"] = True
        dyn["pending"] = list(draws)
        be.counter_add(dyn["it"], set_to=0)

    def _graph_rewind(self):
        ops.backend().counter_add(self._dim_dyn["it"], set_to=0)

    def _graph_step(self):
        ops.backend().counter_add(self._dim_dyn["it"], delta=1)

    def _graph_end(self, ok):
        self._dim_dyn["active"] = False
        if ok:
            self._dim_dyn["pending"] = None
