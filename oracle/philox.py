"""TEST INFRASTRUCTURE (oracle): numpy restatement of how torch's CUDA `Tensor.uniform_(from, to)` fills a contiguous fp32
tensor — the draw the reference makes per VMI-FGSM neighbour (transferattack/gradient/vmifgsm.py:50,
`torch.zeros_like(delta).uniform_(-beta*eps, beta*eps)`) — so that `ta_neighbor_stage_philox` can generate the same numbers
inside the staging kernel. Followed, not copied: ATen/native/cuda/DistributionTemplates.h (calc_execution_policy,
distribution_elementwise_grid_stride_kernel, uniform_kernel) of the installed torch 2.11 and cuRAND's Philox4_32_10
(curand_kernel.h: curand_init / skipahead / curand4; curand_uniform.h: x * 2^-32 + 2^-33); Philox4x32-10 itself is the
published Random123 algorithm (Salmon et al., SC'11) and is pinned below against its known-answer vectors.

  element li  ->  thread idx = li % T, draw q = li // T, call j = q // 4, lane ii = q % 4         (T = 256 * grid)
  counter     =  (offset // 4 + j  [low 64 bits],  idx [high 64 bits]),   key = seed
  value       =  fma(u, to - from, from) with u = float(x_ii) * 2^-32 + 2^-33;  value == to -> from
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """vectorised over uint64 arrays holding 32-bit values"""
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0 = np.uint64(k0) & MASK; k1 = np.uint64(k1) & MASK
    for r in range(10):
        p0 = M0 * c0; p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        if r < 9:
            k0 = (k0 + np.uint64(W0)) & MASK; k1 = (k1 + np.uint64(W1)) & MASK
    return c0, c1, c2, c3


def torch_uniform_policy(numel, sm_count, max_threads_per_sm):
    """(total threads T, philox offset increment) of calc_execution_policy(numel, unroll_factor=4)"""
    block = 256
    grid = min(sm_count * (max_threads_per_sm // block), (numel + block - 1) // block)
    T = block * grid
    return T, ((numel - 1) // (T * 4) + 1) * 4


def torch_uniform(numel, seed, offset, frm, to, T, fma=True):
    """the fp32 values torch's CUDA uniform_(frm, to) writes into a contiguous tensor of `numel` elements"""
    li = np.arange(numel, dtype=np.uint64)
    idx = li % np.uint64(T); q = li // np.uint64(T)
    j = q // np.uint64(4); ii = (q % np.uint64(4)).astype(np.int64)
    ctr = np.uint64(offset // 4) + j
    out = philox4x32_10(ctr & MASK, ctr >> np.uint64(32), idx & MASK, idx >> np.uint64(32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    x = np.choose(ii, [o.astype(np.uint32) for o in out])
    u = x.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)       # exact scaling, one rounding in the add
    f, t = np.float32(frm), np.float32(to)
    rng = np.float32(t - f)
    if fma:
        v = (u.astype(np.float64) * np.float64(rng) + np.float64(f)).astype(np.float32)   # 24x24-bit product exact in f64; one rounding
    else:
        v = (u * rng).astype(np.float32) + f
    return np.where(v == t, f, v).astype(np.float32)
