"""Multi-GPU execution of the hot path: one process per GPU (``torch.distributed``, NCCL over NVLink on the box, gloo in
the CPU tests). The reference is single-GPU (``main.py:31`` pins one device; ``EnsembleModel`` runs its members one
after another on it, utils.py:94-97), so everything here is new functionality behind the same plugin API.

1. Batch sharding — ``shard`` / ``run_sharded``.  Samples are independent on this path (eval-mode BN, per-sample
   L1-normalisation / epsilon-ball / box clamp), so each rank attacks a contiguous slice of the batch and there is NO
   collective on the data path. Two details keep a sharded run equal to the unsharded one shard by shard:
   * ``CrossEntropyLoss`` is a batch MEAN (attack.py:160): a shard's gradient is the full-batch gradient times
     B/B_shard, a factor that cancels exactly in ``g / mean|g|`` when it is a power of two and at rounding level otherwise;
   * DIM draws ONE (coin, size, pad) per call from the CPU generator (dim.py:47-62): ``sync_host_rng`` seeds every
     rank's generators identically so all shards see the transform an unsharded run would.
   Admix mixes images ACROSS the batch (admix.py:44) and therefore runs as replicas, not shards (``run_sharded`` refuses).

2. One surrogate per GPU — ``ShardedEnsembleModel``.  ENS's loss is CE(mean_k logits_k) (utils.py:98-100, attack.py:115).
   With member k on rank k and every rank holding the same (data, delta):
       forward : all_reduce(SUM) of the local logits [B, classes], divided by K
       backward: d mean / d logits_k = gout / K locally, then all_reduce(SUM) of the input gradient [B,3,H,W]
   after which every rank runs the same fused update on the same numbers (replicated, deterministic). A gradient-only
   all-reduce would optimise mean_k CE(logits_k) — a different loss — hence the two collectives.
   With K = 2 the result is bit-identical to the single-device ``EnsembleModel`` (two-term sums commute); for K > 2 the
   collective's summation order differs from ``torch.mean(stack)``'s at rounding level.
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def sync_host_rng(seed, group=None):
    """Give every rank the same torch-CPU / numpy / python generator state (rank 0's `seed` wins)."""
    import random
    rank, world = _world(group)
    if world > 1:
        box = [int(seed)]
        dist.broadcast_object_list(box, src=0, group=group)
        seed = box[0]
    cpu_state = torch.random.default_generator.manual_seed(int(seed))   # CPU generator only: device generators stay per rank
    np.random.seed(int(seed) % (2 ** 32))
    random.seed(int(seed))
    return cpu_state


def shard_bounds(n, rank, world):
    """Contiguous slice [lo, hi) of n samples for `rank`; the first n % world ranks get one extra sample."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(data, label, group=None):
    """This rank's slice of (data, label); label may be [N] or the targeted [2, N] layout (attack.py:76-78)."""
    rank, world = _world(group)
    lo, hi = shard_bounds(data.shape[0], rank, world)
    lab = label[:, lo:hi] if (label.dim() == 2 and label.shape[0] == 2 and label.shape[1] == data.shape[0]) else label[lo:hi]
    return data[lo:hi], lab


def run_sharded(attacker, data, label, seed=None, gather=False, group=None, **kwargs):
    """Attack `data` with the batch sharded over the ranks of `group`. Returns this rank's perturbation slice, or with
    ``gather=True`` the full perturbation on every rank (an all_gather AFTER the attack; not on the data path)."""
    from .input_transformation.admix import Admix
    rank, world = _world(group)
    if isinstance(attacker, Admix) and world > 1:
        raise RuntimeError("Admix mixes images across the batch (admix.py:44): run it as replicas, not batch shards")
    if seed is not None:
        sync_host_rng(seed, group)
    d_local, l_local = shard(data, label, group)
    if d_local.shape[0] == 0:
        delta = torch.empty((0,) + tuple(data.shape[1:]), dtype=torch.float32, device=attacker.device)
    else:
        delta = attacker(d_local, l_local, **kwargs)
    if not gather or world == 1:
        return delta
    sizes = [hi - lo for lo, hi in (shard_bounds(data.shape[0], r, world) for r in range(world))]
    widest = max(sizes)                         # all_gather wants equal shapes: pad the short shards, trim after
    padded = torch.zeros((widest,) + tuple(data.shape[1:]), dtype=delta.dtype, device=delta.device)
    padded[:delta.shape[0]] = delta
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


@torch.no_grad()
def sharded_asr(model, loader, is_targeted, device, group=None, dtype=None):
    """Attack success rate in percent (reference main.py:80-94) with the BATCHES dealt round-robin over the ranks of `group`
    (evaluation is pure inference on independent images → batch-sharded, no data-path collective; two scalars
    are all-reduced at the end). Every rank returns the same number — the one a single-process pass over `loader` returns.
    `dtype` (e.g. torch.bfloat16) runs the victim under autocast: inference only, optional, off by default."""
    rank, world = _world(group)
    counts = torch.zeros(2, dtype=torch.float64, device=device)
    for i, (images, labels, _) in enumerate(loader):
        if i % world != rank:
            continue
        if is_targeted:
            labels = labels[1]
        x = images.to(device, non_blocking=True)
        if dtype is not None and torch.device(device).type == "cuda":
            with torch.autocast("cuda", dtype=dtype):
                pred = model(x).argmax(dim=1)
        else:
            pred = model(x).argmax(dim=1)
        counts[0] += (labels.to(pred.device) == pred).sum()
        counts[1] += labels.shape[0]
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    correct, total = float(counts[0]), float(counts[1])
    return (correct / total) * 100 if is_targeted else (1 - correct / total) * 100


class _SumGradAcrossRanks(torch.autograd.Function):
    """identity forward; backward all_reduce(SUM)s the gradient wrt the (replicated) model input"""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, gout):
        g = gout.contiguous()
        if g.data_ptr() == gout.data_ptr():
            g = g.clone()                      # never reduce in place into autograd's buffer
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


class _MeanLogitsAcrossRanks(torch.autograd.Function):
    """forward: (sum_k logits_k) / K via all_reduce(SUM); backward: gout / K (the local member's share of the mean)"""

    @staticmethod
    def forward(ctx, logits, group, K):
        ctx.K = K
        out = logits.detach().clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out / K

    @staticmethod
    def backward(ctx, gout):
        return gout / ctx.K, None, None


class ShardedEnsembleModel(nn.Module):
    """Ensemble with ONE member per rank: the drop-in for ``EnsembleModel`` (utils.py:82-105, mode 'mean') when the
    attack runs as one process per GPU. ``member`` is this rank's wrapped surrogate; all ranks must feed the same input."""

    def __init__(self, member, group=None):
        super().__init__()
        self.member = member
        self.group = group
        self.device = next(member.parameters()).device
        self.models = [member]                 # this rank's view (attacks that index all members need EnsembleModel)
        self.num_models = _world(group)[1]
        self.mode = 'mean'
        self.type_name = 'ensemble'
        self.softmax = torch.nn.Softmax(dim=1)

    def forward(self, x):
        if self.num_models == 1:
            return self.member(x)
        x = _SumGradAcrossRanks.apply(x, self.group)
        return _MeanLogitsAcrossRanks.apply(self.member(x), self.group, self.num_models)


def make_ens_attack(attack_cls, member, group=None, **kwargs):
    """Instantiate an ENS-style attack class whose surrogate is the per-rank sharded ensemble (``load_model`` is the
    reference's documented override point, attack.py:40-65). The base ``Attack`` treats ``ShardedEnsembleModel`` like any
    module; ``device`` is taken from the member."""
    model = ShardedEnsembleModel(member, group)

    def init_delta(self, data, **kw):
        # every rank must start from the SAME delta (the replicated update assumes it): a random start is drawn from each
        # rank's own device generator (attack.py:131-141), so rank 0's draw is broadcast
        delta = attack_cls.init_delta(self, data, **kw)
        if self.random_start and _world(group)[1] > 1:
            with torch.no_grad():
                grp = group if group is not None else dist.group.WORLD
                dist.broadcast(delta, src=dist.get_global_rank(grp, 0), group=group)
        return delta

    # collectives inside forward/backward: keep the loop eager (NCCL inside a captured graph is not exercised here)
    P = type("Sharded" + attack_cls.__name__, (attack_cls,),
             {"load_model": lambda self, _n: model, "use_cuda_graph": False, "init_delta": init_delta})
    return P(model_name="sharded-ensemble", device=model.device, **kwargs)


# =====================================================================================================================
# 3. ENS with the collective FUSED into the update: one kernel does reduce-scatter + update + all-gather over NVLink
# =====================================================================================================================
class _GatherMeanLogits(torch.autograd.Function):
    """forward: all_gather the K members' logits and average them with the reference's own ops — torch.mean(torch.stack(...))
    (utils.py:97-99) — so the mean is bit-identical to the single-device EnsembleModel for any K; backward: gout / K."""

    @staticmethod
    def forward(ctx, logits, group, K):
        ctx.K = K
        parts = [torch.empty_like(logits) for _ in range(K)]
        dist.all_gather(parts, logits.detach().contiguous(), group=group)
        return torch.mean(torch.stack(parts, dim=0), dim=0)

    @staticmethod
    def backward(ctx, gout):
        return gout / ctx.K, None, None


def _symm():
    import torch.distributed._symmetric_memory as symm_mem
    return symm_mem


class FusedP2PEnsembleLoop:
    """The ENS iteration with one surrogate per GPU where the gradient exchange is fused into the update kernel
    (``ta_fused_allreduce_update_linf``): every rank forwards/backwards the FULL batch through its own member, publishes its
    input gradient in symmetric (peer-mapped) memory, and then updates only the samples it OWNS — reading the K gradients of
    those samples straight from the peers over NVLink, and writing the next model input into every peer's buffer.
    Per GPU and iteration that is (K-1)/K * |g| in and (K-1)/K * |x| out over NVLink — the volume of reduce-scatter +
    all-gather — with no separate collective kernels and no replicated update. `mean|g|` is formed in the kernel
    (mean_mode 'exact'); the gradient sum follows autograd's accumulation order, the logits mean uses the reference's own
    ops, so for any K the result equals the single-device EnsembleModel run with mean_mode='exact' bit for bit."""

    def __init__(self, attacker, group=None):
        self.atk = attacker
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.K = dist.get_rank(self.group), dist.get_world_size(self.group)
        self._state = {}

    def _buffers(self, data):
        key = (tuple(data.shape), str(data.device))
        st = self._state.get(key)
        if st is not None:
            return st
        symm_mem = _symm()
        G = symm_mem.empty(*data.shape, dtype=torch.float32, device=data.device)
        X = symm_mem.empty(*data.shape, dtype=torch.float32, device=data.device)
        hg = symm_mem.rendezvous(G, self.group)
        hx = symm_mem.rendezvous(X, self.group)
        import ctypes
        st = {"G": G, "X": X, "hg": hg, "hx": hx,
              "g_ptrs": (ctypes.c_void_p * self.K)(*[int(p) for p in hg.buffer_ptrs]),
              "x_ptrs": (ctypes.c_void_p * self.K)(*[int(p) for p in hx.buffer_ptrs]),
              "m": torch.zeros_like(data), "scale_out": torch.zeros(data.shape[0], device=data.device, dtype=torch.float32)}
        self._state[key] = st
        return st

    def __call__(self, data, label):
        from . import _lib, ops
        from .utils import img_max, img_min
        atk = self.atk
        atk.model.eval()
        if atk.targeted:
            assert len(label) == 2
            label = label[1]
        if atk.norm != 'linfty':
            raise RuntimeError("the fused P2P ensemble loop implements the L-inf update")
        data = atk._to_device(data).contiguous()
        label = atk._to_device(label)
        B = data.shape[0]
        n = data.numel() // B
        lo, hi = shard_bounds(B, self.rank, self.K)
        st = self._buffers(data)
        be = ops.backend()
        lib = _lib.load()
        delta = atk.init_delta(data).detach()
        if atk.random_start:
            dist.broadcast(delta, src=dist.get_global_rank(self.group, 0), group=self.group)      # one draw for all ranks
        st["m"].zero_()
        be.stage_add(data, delta, out=st["X"])
        st["hx"].barrier(channel=0)
        stream = torch.cuda.current_stream(data.device)
        for _ in range(atk.epoch):
            x_leaf = st["X"].detach().requires_grad_(True)
            logits = _GatherMeanLogits.apply(atk.get_logits(atk.transform(x_leaf, momentum=0)), self.group, self.K)
            loss = atk.get_loss(logits, label)
            g = torch.autograd.grad(loss, x_leaf)[0]
            st["G"].copy_(g)
            st["hg"].barrier(channel=0)                     # every rank's gradient is published
            _lib.check(lib.ta_fused_allreduce_update_linf(
                st["g_ptrs"], st["x_ptrs"], self.K, st["m"].data_ptr(), st["m"].data_ptr(), delta.data_ptr(), delta.data_ptr(),
                data.data_ptr(), None, st["scale_out"].data_ptr(), _lib.TA_MEAN_EXACT, float(atk.decay), float(atk.alpha),
                float(atk.epsilon), float(img_min), float(img_max), lo, hi - lo, n, stream.cuda_stream), "ta_fused_allreduce_update_linf")
            st["hx"].barrier(channel=0)                     # every owner's x_adv rows are visible everywhere
        # hand every rank the full perturbation (after the attack; not on the per-iteration path)
        sizes = [h - l for l, h in (shard_bounds(B, r, self.K) for r in range(self.K))]
        widest = max(sizes)
        mine = torch.zeros((widest,) + tuple(data.shape[1:]), dtype=torch.float32, device=data.device)
        mine[:hi - lo] = delta[lo:hi]
        parts = [torch.empty_like(mine) for _ in range(self.K)]
        dist.all_gather(parts, mine, group=self.group)
        return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def make_fused_p2p_ens(attack_cls, member, group=None, **kwargs):
    """ENS-style attack (`attack_cls`, e.g. transferattack_b200.ensemble.ens.ENS) with this rank's `member` as its surrogate,
    run through FusedP2PEnsembleLoop. Returns a callable (data, label) → full perturbation."""
    P = type("P2P" + attack_cls.__name__, (attack_cls,), {"load_model": lambda self, _n: member, "use_cuda_graph": False})
    atk = P(model_name="p2p-ensemble-member", device=next(member.parameters()).device, **kwargs)
    atk.mean_mode = 'exact'
    return FusedP2PEnsembleLoop(atk, group)
