"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): restatement of the summation ORDER of torch's CUDA
``x.mean(dim=(1,2,3))`` for a contiguous fp32 [B, C, H, W] tensor — the reference's ``grad.abs().mean(dim=(1,2,3))``
(transferattack/attack.py:128).

The algorithm lives in PyTorch (ATen/native/cuda/Reduce.cuh: ``setReduceConfig``, ``ReduceOp::run`` — thread_reduce with
vt0 = 4 accumulators, block_y_reduce, block_x_reduce, global_reduce; ATen/native/cuda/ReduceMomentKernel.cu: ``MeanOps`` with
factor = float(num_outputs) / numel), which is a dependency of the reference (requirements.txt pins torch), not a file in
/root/reference. This module restates its published launch policy and tree for the case the hot path uses (reduction over
the fastest dimension, more than one output, fp32 in / fp32 accumulate) and is pinned by running it against torch itself on
the GPU box (tools/diag_aten_mean.py → profiles/diag_aten_mean_r2.json; tests/test_kernels_gpu.py).

``config`` mirrors the host policy; ``emulate`` replays the tree with torch ops (one rounding per add) on any device;
``emulate_numpy`` does the same in numpy for CPU tests of the kernels' index logic.
"""
import math

import numpy as np

MAX_NUM_THREADS = 512          # Reduce.cuh: mem::utils / ReduceConfig::MAX_NUM_THREADS for 4-byte types
VT0 = 4                        # gpu_reduce_kernel<scalar_t, out_t, vt0 = 4>
WARP = 32


def _last_pow2(n):
    p = 1
    while p * 2 <= n:
        p *= 2
    return p


def _div_up(a, b):
    return (a + b - 1) // b


def config(B, n, sm_count=148, max_threads_per_sm=2048):
    """ReduceConfig for a [B, n] fp32 tensor reduced over n (stride 1), B >= 2 outputs (setReduceConfig, Reduce.cuh).
    Returns a dict, or None when the launch falls outside the family restated here (then the product keeps ATen's op)."""
    if B < 2 or n < 32:
        return None                      # B == 1: the iterator is 1-D and ATen vectorises the input loads (other tree)
    dim0, dim1 = n, B
    d0p = _last_pow2(dim0) if dim0 < MAX_NUM_THREADS else MAX_NUM_THREADS
    d1p = _last_pow2(dim1) if dim1 < MAX_NUM_THREADS else MAX_NUM_THREADS
    bw = min(d0p, WARP)
    bh = min(d1p, MAX_NUM_THREADS // bw)
    bw = min(d0p, MAX_NUM_THREADS // bh)
    num_threads = bw * bh
    if num_threads != MAX_NUM_THREADS or bw < WARP or bh > 16:
        return None                      # the replay kernels assume ATen's full 512-thread block
    step_input = bw                      # input_mult[0] = split_input(block_width)
    vpt = _div_up(n, step_input)
    if not (vpt >= bh * 16 or vpt >= 256):
        return None                      # each warp row reduces its own output: not the hot path's shape
    step_input *= bh                     # input_mult[1] = split_input(block_height)
    vpt = _div_up(n, step_input)
    grid_x = B                           # one output per block
    blocks_per_sm = max_threads_per_sm // num_threads
    target = sm_count * blocks_per_sm
    cpo = 1
    if vpt >= 256 and grid_x <= target:
        c1 = _div_up(target, grid_x)
        c2 = _div_up(vpt, 16)
        c3 = _div_up(vpt, 256)
        cpo = max(min(c1, c2), c3)
    if cpo > WARP:
        return None                      # the kernels' final tree holds one partial per lane of one warp
    return {"bw": bw, "bh": bh, "cpo": cpo, "threads": num_threads, "stride": num_threads * cpo}


def _tree(v, xp, add, variants):
    """v: [B, cpo, bh, bw] thread values → [B] sums, in ReduceOp::run's order."""
    B, cpo, bh, bw = v.shape
    y_first = variants.get("y_first", True)
    asc = variants.get("shfl_ascending", True)

    def yred(t):                           # block_y_reduce: offsets bh/2 .. 1 through shared memory
        h = t.shape[2] // 2
        while h >= 1:
            t = add(t[:, :, :h], t[:, :, h:2 * h])
            h //= 2
        return t                           # [.., 1, bw]

    def xred(t):                           # block_x_reduce: shared memory down to 32 lanes, then shuffles
        w = t.shape[3]
        off = w // 2
        while off >= WARP:
            t = add(t[..., :off], t[..., off:2 * off])
            off //= 2
        return _shfl(t, xp, add, asc)      # [.., 1]

    if y_first:
        t = xred(yred(v))
    else:
        t = yred(xred(v))
    blk = t.reshape(B, cpo)
    if cpo == 1:
        return blk[:, 0]
    # global_reduce, last block: thread (tx, 0) holds sum of staging[tx], staging[tx + 512], ...; then the same two trees
    lanes = xp.zeros((B, MAX_NUM_THREADS), dtype=blk.dtype) if xp is np else xp.zeros((B, MAX_NUM_THREADS), dtype=blk.dtype, device=blk.device)
    for i in range(cpo):
        lanes[:, i % MAX_NUM_THREADS] = add(lanes[:, i % MAX_NUM_THREADS], blk[:, i])
    t = lanes.reshape(B, 1, bh, bw)
    t = xred(yred(t)) if y_first else yred(xred(t))
    return t.reshape(B)


def _shfl(t, xp, add, ascending):
    """warp shuffle-down tree over the last dim (<= 32 wide, zero-extended to 32): lane 0's value"""
    w = t.shape[-1]
    if w < WARP:
        pad = list(t.shape); pad[-1] = WARP - w
        z = xp.zeros(pad, dtype=t.dtype) if xp is np else xp.zeros(pad, dtype=t.dtype, device=t.device)
        t = xp.concatenate([t, z], -1) if xp is np else xp.cat([t, z], -1)
    offs = [1, 2, 4, 8, 16] if ascending else [16, 8, 4, 2, 1]
    for off in offs:
        # lane l gets value[l] + value[l + off] (out of range: its own value — never reaches lane 0)
        if xp is np:
            sh = np.concatenate([t[..., off:], t[..., WARP - off:]], -1)
        else:
            sh = xp.cat([t[..., off:], t[..., WARP - off:]], -1)
        t = add(t, sh)
    return t[..., :1]


def _emulate(x, cfg, xp, variants):
    B, n = x.shape
    S = cfg["stride"]
    J = _div_up(n, S)
    if xp is np:
        xpad = np.zeros((B, J * S), np.float32); xpad[:, :n] = x
        add = lambda a, b: (a.astype(np.float32) + b.astype(np.float32)).astype(np.float32)
        zeros = lambda: np.zeros((B, S), np.float32)
    else:
        xpad = xp.zeros((B, J * S), dtype=x.dtype, device=x.device); xpad[:, :n] = x
        add = lambda a, b: a + b
        zeros = lambda: xp.zeros((B, S), dtype=x.dtype, device=x.device)
    X = xpad.reshape(B, J, S)
    acc = [zeros() for _ in range(VT0)]
    for j in range(J):                     # thread_reduce_impl: element j of a thread goes to accumulator j % vt0 (tail included)
        acc[j % VT0] = add(acc[j % VT0], X[:, j])
    v = add(add(add(acc[0], acc[1]), acc[2]), acc[3])
    v = v.reshape(B, cfg["cpo"], cfg["bh"], cfg["bw"])
    s = _tree(v, xp, add, variants)
    if variants.get("mul_factor", True):
        if xp is np:
            factor = np.float32(np.float32(B) / np.float32(B * n))
            return (s * factor).astype(np.float32)
        factor = float(np.float32(np.float32(B) / np.float32(B * n)))
        return s * factor
    return s / n


def emulate(x, sm_count=148, max_threads_per_sm=2048, **variants):
    """x: torch [B, n] fp32 (already |g|) on any device → [B] means in ATen's CUDA order; None if outside the family."""
    import torch
    cfg = config(x.shape[0], x.shape[1], sm_count, max_threads_per_sm)
    return None if cfg is None else _emulate(x, cfg, torch, variants)


def emulate_numpy(x, sm_count=148, max_threads_per_sm=2048, **variants):
    cfg = config(x.shape[0], x.shape[1], sm_count, max_threads_per_sm)
    return None if cfg is None else _emulate(np.asarray(x, np.float32), cfg, np, variants)
