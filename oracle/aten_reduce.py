"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): restatement of the summation ORDER of torch's CUDA
``x.mean(dim=(1,2,3))`` for a contiguous fp32 [B, C, H, W] tensor — the reference's ``grad.abs().mean(dim=(1,2,3))``
(transferattack/attack.py:128).

The algorithm lives in PyTorch, a dependency of the reference (requirements.txt pins torch), not in /root/reference. The
installed build (torch 2.11.0+cu128) ships the source it was compiled from as a header:
``torch/include/ATen/native/cuda/Reduce.cuh`` — line numbers below refer to it — plus ``ATen/native/SharedReduceOps.h:165-192``
(``MeanOps``: reduce = combine = a + b, project = a * factor). What is restated (fp32 in, fp32 accumulate, vt0 = 4,
input_vec_size = 4, reduction over the fastest dimension, n % 4 == 0 and 16-byte aligned rows so that there is no head/tail):

  setReduceConfig (Reduce.cuh:1033-1178)  "vectorize along input" (one reduced dimension after coalescing, n >= 128): dim0 = n/4
        vectors; block (bw, bh) from set_block_dimension (:100-108); lanes split the input (input_mult[0]); warps split it too
        when values_per_thread >= min(16*bh, 256); ctas_per_output = max(min(ceil(target/B), ceil(vpt/16)), ceil(vpt/256))
        when vpt >= 256 and B <= SMs * (max threads per SM / block threads);
  input_vectorized_thread_reduce_impl (:500-559)  virtual thread t = tx + bw*ty + (bw*bh)*cta loads the 128-bit vectors
        t, t+S, t+2S, ... (S = bw*bh*ctas_per_output) and adds component i into accumulator i, in order; value = ((a0+a1)+a2)+a3;
  block_x_reduce (:634-672)  FIRST: shared-memory tree over tx down to 32 lanes (offsets bw/2 .. 32), then shfl_down with
        DEcreasing offsets 16, 8, 4, 2, 1;
  block_y_reduce (:674-692)  then the shared-memory tree over ty (offsets bh/2 .. 1);
  global_reduce (:787-876)  the last CTA: thread with linear id i takes staging[i] (i < ctas_per_output), then block_y_reduce,
        then block_x_reduce;
  project  sum * factor, factor = (float)B / (float)(B*n)  (ReduceMomentKernel.cu; confirmed by tools/diag_aten_mean.py).

Pinned by running it against torch itself on the GPU box (tools/diag_aten_mean.py → profiles/diag_aten_mean_r2.json;
tests/test_kernels_gpu.py). ``config`` mirrors the host policy; ``emulate`` replays the tree with torch ops (one rounding per
add) on any device; ``emulate_numpy`` does the same in numpy for CPU tests of the kernels' index logic.
"""
import numpy as np

MAX_NUM_THREADS = 512          # Reduce.cuh:62 mnt_wrapper<float>::MAX_NUM_THREADS
VEC = 4                        # input_vec_size = vt0 = 4 (gpu_reduce_kernel's defaults, Reduce.cuh:1186)
WARP = 32


def _last_pow2(n):
    p = 1
    while p * 2 <= n:
        p *= 2
    return p


def _div_up(a, b):
    return (a + b - 1) // b


def config(B, n, sm_count=148, max_threads_per_sm=2048):
    """ReduceConfig for a contiguous [B, n] fp32 tensor reduced over n (setReduceConfig, Reduce.cuh:1033-1178).
    Returns a dict, or None when the launch falls outside the family restated here (then the product keeps ATen's op)."""
    if B < 1 or n < 128 or n % VEC != 0:
        return None                      # n < 128: not vectorised; n % 4: head/tail elements take another path (:505-553)
    dim0, dim1 = n // VEC, B
    d0p = _last_pow2(dim0) if dim0 < MAX_NUM_THREADS else MAX_NUM_THREADS
    d1p = _last_pow2(dim1) if dim1 < MAX_NUM_THREADS else MAX_NUM_THREADS
    bw = min(d0p, WARP)
    bh = min(d1p, MAX_NUM_THREADS // bw)
    bw = min(d0p, MAX_NUM_THREADS // bh)
    nt = bw * bh
    if bw < WARP or bh > 16:
        return None
    step = bw                            # input_mult[0] = split_input(block_width)
    vpt = _div_up(n, step)               # values_per_thread(): num_inputs counts ELEMENTS, the steps count vectors (as in the source)
    if vpt < min(bh * 16, 256):
        return None                      # each warp row reduces its own output (output_mult[1]): not the hot path's shape
    step *= bh                           # input_mult[1] = split_input(block_height)
    vpt = _div_up(n, step)
    target = sm_count * (max_threads_per_sm // nt)
    cpo = 1
    if vpt >= 256 and B <= target:
        cpo = max(min(_div_up(target, B), _div_up(vpt, 16)), _div_up(vpt, 256))
    if cpo > bw:
        return None                      # the kernels' final tree holds one partial per x position of one block row
    return {"bw": bw, "bh": bh, "cpo": cpo, "threads": nt, "stride": nt * cpo}


def _zeros(xp, shape, like):
    return np.zeros(shape, np.float32) if xp is np else xp.zeros(shape, dtype=like.dtype, device=like.device)


def _cat(xp, parts):
    return np.concatenate(parts, -1) if xp is np else xp.cat(parts, -1)


def _xred(t, xp, add, descending=True):
    """block_x_reduce over the last dim: shared-memory levels down to 32 lanes, then shfl_down (offsets 16..1); [..., 1]"""
    off = t.shape[-1] // 2
    while off >= WARP:
        t = add(t[..., :off], t[..., off:2 * off])
        off //= 2
    w = t.shape[-1]
    if w < WARP:
        pad = list(t.shape); pad[-1] = WARP - w
        t = _cat(xp, [t, _zeros(xp, pad, t)])
    for o in ([16, 8, 4, 2, 1] if descending else [1, 2, 4, 8, 16]):
        t = add(t, _cat(xp, [t[..., o:], t[..., WARP - o:]]))       # lane l += lane l+o (out of range: itself; never reaches lane 0)
    return t[..., :1]


def _yred(t, add):
    """block_y_reduce over dim -2: offsets bh/2 .. 1; [..., 1, w]"""
    h = t.shape[-2] // 2
    while h >= 1:
        t = add(t[..., :h, :], t[..., h:2 * h, :])
        h //= 2
    return t


def _emulate(x, cfg, xp, variants):
    B, n = x.shape
    bw, bh, cpo, S = cfg["bw"], cfg["bh"], cfg["cpo"], cfg["stride"]
    nvec = n // VEC
    J = _div_up(nvec, S)
    if xp is np:
        add = lambda a, b: (a.astype(np.float32) + b.astype(np.float32)).astype(np.float32)
    else:
        add = lambda a, b: a + b
    xpad = _zeros(xp, (B, J * S * VEC), x)
    xpad[:, :n] = x                        # rows past the end contribute +0.0f (exact); ATen skips them
    X = xpad.reshape(B, J, S, VEC)
    acc = _zeros(xp, (B, S, VEC), x)
    for j in range(J):                     # component i of every vector of the thread goes to accumulator i, vectors in order
        acc = add(acc, X[:, j])
    v = add(add(add(acc[..., 0], acc[..., 1]), acc[..., 2]), acc[..., 3])
    v = v.reshape(B, cpo, bh, bw)
    desc = variants.get("shfl_descending", True)
    if variants.get("x_first", True):
        blk = _yred(_xred(v, xp, add, desc), add)
    else:
        blk = _xred(_yred(v, add), xp, add, desc)
    blk = blk.reshape(B, cpo)
    if cpo == 1:
        s = blk[:, 0]
    else:                                  # global_reduce's last block: partial i at linear thread id i; y tree, then x tree
        lanes = _zeros(xp, (B, bh * bw), x)
        lanes[:, :cpo] = blk
        s = _xred(_yred(lanes.reshape(B, 1, bh, bw), add), xp, add, desc).reshape(B)
    if variants.get("mul_factor", True):
        factor = np.float32(np.float32(B) / np.float32(B * n))
        return (s * factor).astype(np.float32) if xp is np else s * float(factor)
    return (s / np.float32(n)).astype(np.float32) if xp is np else s / n


def emulate(x, sm_count=148, max_threads_per_sm=2048, **variants):
    """x: torch [B, n] fp32 (already |g|) on any device → [B] means in ATen's CUDA order; None if outside the family."""
    import torch
    cfg = config(x.shape[0], x.shape[1], sm_count, max_threads_per_sm)
    return None if cfg is None else _emulate(x, cfg, torch, variants)


def emulate_numpy(x, sm_count=148, max_threads_per_sm=2048, **variants):
    x = np.asarray(x, np.float32)
    cfg = config(x.shape[0], x.shape[1], sm_count, max_threads_per_sm)
    return None if cfg is None else _emulate(x, cfg, np, variants)
