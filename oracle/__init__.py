"""TEST INFRASTRUCTURE ONLY — the CPU oracle for the TransferAttack hot path.

Two pieces:
  * ``ta_oracle.c``  — plain-C restatement of every per-iteration op (numpy arrays in/out via ctypes),
    the checker for the CUDA kernels;
  * ``torch_ref.py`` — eager-PyTorch restatement of the reference ``Attack`` loop and the in-scope
    plugin classes, used as the end-to-end comparator (same device, same surrogate) and as the CPU
    baseline timed by ``bench.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this package.  ``transferattack_b200`` never does: the product path fails loudly when its
CUDA library is missing instead of falling back to anything here.

Parity pin: against outputs of the unmodified reference (``tests/golden/make_golden.py`` →
``tests/golden/*.npz``), because the reference ships no tests or golden vectors.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libta_oracle.so")
_SRC = os.path.join(_HERE, "ta_oracle.c")
_lib = None


def build(force=False):
    """Compile ta_oracle.c → libta_oracle.so (gcc, no FMA contraction)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-fvisibility=hidden", "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


_F = ctypes.POINTER(ctypes.c_float)


def _fp(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_F)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def abs_mean_per_sample(g):
    g = _c(g); B = g.shape[0]; n = g.size // B
    out = np.empty(B, np.float32)
    lib().orc_abs_mean_per_sample(_fp(g), _fp(out), B, ctypes.c_int64(n))
    return out


def momentum(g, m, scale, decay):
    g = _c(g); m = _c(m); scale = _c(scale); B = g.shape[0]; n = g.size // B
    out = np.empty_like(g)
    lib().orc_momentum(_fp(g), _fp(m), _fp(scale), ctypes.c_float(decay), _fp(out), B, ctypes.c_int64(n))
    return out


def update_linf(delta, data, direction, alpha, eps, lo=0.0, hi=1.0, alpha_t=None, dir_mode=0):
    delta = _c(delta); data = _c(data); direction = _c(direction); alpha_t = _c(alpha_t)
    out = np.empty_like(delta)
    lib().orc_update_linf(_fp(delta), _fp(data), _fp(direction), _fp(alpha_t), ctypes.c_float(alpha),
                          ctypes.c_float(eps), ctypes.c_float(lo), ctypes.c_float(hi), int(dir_mode), _fp(out),
                          ctypes.c_int64(delta.size))
    return out


def clamp_box(delta, data, lo=0.0, hi=1.0):
    delta = _c(delta); data = _c(data)
    out = np.empty_like(delta)
    lib().orc_clamp_box(_fp(delta), _fp(data), ctypes.c_float(lo), ctypes.c_float(hi), _fp(out),
                        ctypes.c_int64(delta.size))
    return out


def update_l2(delta, data, g, alpha, eps, lo=0.0, hi=1.0):
    delta = _c(delta); data = _c(data); g = _c(g); B = g.shape[0]; n = g.size // B
    out = np.empty_like(delta)
    lib().orc_update_l2(_fp(delta), _fp(data), _fp(g), ctypes.c_float(alpha), ctypes.c_float(eps),
                        ctypes.c_float(lo), ctypes.c_float(hi), _fp(out), B, ctypes.c_int64(n))
    return out


def init_l2_scale(delta, r, data, eps, lo=0.0, hi=1.0):
    delta = _c(delta); r = _c(r); data = _c(data); B = delta.shape[0]; n = delta.size // B
    out = np.empty_like(delta)
    lib().orc_init_l2_scale(_fp(delta), _fp(r), _fp(data), ctypes.c_float(eps), ctypes.c_float(lo),
                            ctypes.c_float(hi), _fp(out), B, ctypes.c_int64(n))
    return out


def fused_update_linf(g, m, delta, data, scale, decay, alpha, eps, lo=0.0, hi=1.0, want_xadv=True):
    g = _c(g); m = _c(m); delta = _c(delta); data = _c(data); scale = _c(scale)
    B = g.shape[0]; n = g.size // B
    m_out = np.empty_like(g); d_out = np.empty_like(g); x_out = np.empty_like(g) if want_xadv else None
    lib().orc_fused_update_linf(_fp(g), _fp(m), _fp(m_out), _fp(delta), _fp(d_out), _fp(data), _fp(x_out),
                                _fp(scale), ctypes.c_float(decay), ctypes.c_float(alpha), ctypes.c_float(eps),
                                ctypes.c_float(lo), ctypes.c_float(hi), B, ctypes.c_int64(n))
    return m_out, d_out, x_out


def fused_update_linf_nf(g, m, delta, data, scale, decay, alpha, eps, mean, std, grad_wrt_xn, lo=0.0, hi=1.0):
    """Normalize folded into the fused tail restated as the chain of reference ops it replaces
    (utils.py:72-79 around attack.py:88,124-153): Normalize's adjoint g / std (when the gradient is w.r.t. the normalised
    input), the per-sample mean of |g| (when `scale` is None), the unfused tail, then Normalize's forward on data + delta'."""
    g = _c(g)
    if grad_wrt_xn:
        g = normalize_bwd(g, std)
    if scale is None:
        scale = abs_mean_per_sample(g)
    m_out, d_out, x_out = fused_update_linf(g, m, delta, data, scale, decay, alpha, eps, lo, hi, want_xadv=True)
    return m_out, d_out, normalize_fwd(x_out, mean, std), _c(scale)


def stage_add(data, delta, look=None, coef=0.0):
    data = _c(data); delta = _c(delta); look = _c(look)
    out = np.empty_like(data)
    lib().orc_stage_add(_fp(data), _fp(delta), _fp(look), ctypes.c_float(coef), _fp(out), ctypes.c_int64(data.size))
    return out


def normalize_fwd(x, mean, std):
    x = _c(x); mean = _c(mean); std = _c(std); B, C = x.shape[:2]; plane = x.size // (B * C)
    out = np.empty_like(x)
    lib().orc_normalize_fwd(_fp(x), _fp(mean), _fp(std), _fp(out), B, C, ctypes.c_int64(plane))
    return out


def normalize_bwd(gout, std):
    gout = _c(gout); std = _c(std); B, C = gout.shape[:2]; plane = gout.size // (B * C)
    out = np.empty_like(gout)
    lib().orc_normalize_bwd(_fp(gout), _fp(std), _fp(out), B, C, ctypes.c_int64(plane))
    return out


def sim_fwd(x, S):
    x = _c(x)
    out = np.empty((S * x.shape[0],) + x.shape[1:], np.float32)
    lib().orc_sim_fwd(_fp(x), _fp(out), S, ctypes.c_int64(x.size))
    return out


def sim_bwd(gout, S):
    gout = _c(gout)
    gin = np.empty((gout.shape[0] // S,) + gout.shape[1:], np.float32)
    lib().orc_sim_bwd(_fp(gout), _fp(gin), S, ctypes.c_int64(gin.size))
    return gin


def admix_fwd(x, perm, strength, S):
    x = _c(x); perm = np.ascontiguousarray(perm, dtype=np.int32); A, B = perm.shape; n = x.size // B
    out = np.empty((S * A * B,) + x.shape[1:], np.float32)
    lib().orc_admix_fwd(_fp(x), perm.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_float(strength),
                        _fp(out), S, A, B, ctypes.c_int64(n))
    return out


def admix_bwd(gout, S, A):
    gout = _c(gout); B = gout.shape[0] // (S * A); n = gout.size // gout.shape[0]
    gin = np.empty((B,) + gout.shape[1:], np.float32)
    lib().orc_admix_bwd(_fp(gout), _fp(gin), S, A, B, ctypes.c_int64(n))
    return gin


def dim_fwd(x, rnd, R, pad_top, pad_left, blend=0):
    x = _c(x); S = x.shape[-1]; planes = x.size // (S * S)
    out = np.empty_like(x)
    lib().orc_set_dim_blend(int(blend))
    rc = lib().orc_dim_fwd(_fp(x), _fp(out), planes, S, rnd, R, pad_top, pad_left)
    lib().orc_set_dim_blend(0)
    assert rc == 0
    return out


def dim_bwd(gout, rnd, R, pad_top, pad_left):
    gout = _c(gout); S = gout.shape[-1]; planes = gout.size // (S * S)
    gin = np.empty_like(gout)
    rc = lib().orc_dim_bwd(_fp(gout), _fp(gin), planes, S, rnd, R, pad_top, pad_left)
    assert rc == 0
    return gin


def dwconv2d(g, k):
    g = _c(g); k = _c(k); B, C, H, W = g.shape; ks = k.shape[-1]
    out = np.empty_like(g)
    lib().orc_dwconv2d(_fp(g), _fp(k.reshape(C, ks, ks)), ks, _fp(out), B, C, H, W)
    return out


def dwconv2d_sep(g, kcol, krow):
    g = _c(g); kcol = _c(kcol); krow = _c(krow); B, C, H, W = g.shape; ks = kcol.shape[-1]
    out = np.empty_like(g)
    rc = lib().orc_dwconv2d_sep(_fp(g), _fp(kcol), _fp(krow), ks, _fp(out), B, C, H, W)
    assert rc == 0
    return out


def lin_sample_fwd(x, gbar, coef):
    x = _c(x); gbar = _c(gbar); coef = _c(coef); K = coef.size
    out = np.empty((K * x.shape[0],) + x.shape[1:], np.float32)
    lib().orc_lin_sample_fwd(_fp(x), _fp(gbar), _fp(coef), K, _fp(out), ctypes.c_int64(x.size))
    return out


def lin_sample_bwd(gout, K):
    gout = _c(gout)
    gin = np.empty((gout.shape[0] // K,) + gout.shape[1:], np.float32)
    lib().orc_lin_sample_bwd(_fp(gout), _fp(gin), K, ctypes.c_int64(gin.size))
    return gin


def neighbor_stage(data, delta, noise, look=None, coef=0.0):
    data = _c(data); delta = _c(delta); noise = _c(noise); look = _c(look)
    out = np.empty_like(data)
    lib().orc_neighbor_stage(_fp(data), _fp(delta), _fp(noise), _fp(look), ctypes.c_float(coef), _fp(out),
                             ctypes.c_int64(data.size))
    return out


def accumulate(acc, g, first):
    g = _c(g)
    acc = np.array(acc, dtype=np.float32, copy=True) if acc is not None else np.empty_like(g)
    lib().orc_accumulate(_fp(acc), _fp(g), int(bool(first)), ctypes.c_int64(g.size))
    return acc


def variance_finalize(acc, cur, num_neighbor):
    acc = _c(acc); cur = _c(cur)
    out = np.empty_like(acc)
    lib().orc_variance_finalize(_fp(acc), _fp(cur), int(num_neighbor), _fp(out), ctypes.c_int64(acc.size))
    return out


def add(a, b):
    a = _c(a); b = _c(b)
    out = np.empty_like(a)
    lib().orc_add(_fp(a), _fp(b), _fp(out), ctypes.c_int64(a.size))
    return out


def quantize_u8(data, delta, to_nhwc=True):
    data = _c(data); delta = _c(delta); B, C = data.shape[:2]; plane = data.size // (B * C)
    shape = (B,) + tuple(data.shape[2:]) + (C,) if to_nhwc else data.shape
    out = np.empty(shape, np.uint8)
    lib().orc_quantize_u8(_fp(data), _fp(delta), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), B, C,
                          ctypes.c_int64(plane), int(bool(to_nhwc)))
    return out


def pi_cut_noise(amp, m, coef, eps):
    m = _c(m); amp = _c(amp)
    amp_out = np.empty_like(m); cut = np.empty_like(m)
    lib().orc_pi_cut_noise(_fp(amp), _fp(m), ctypes.c_float(coef), ctypes.c_float(eps), _fp(amp_out), _fp(cut), ctypes.c_int64(m.size))
    return amp_out, cut


def pi_update_linf(delta, data, g, conv, amp, alpha, gamma, eps, lo=0.0, hi=1.0):
    delta = _c(delta); data = _c(data); g = _c(g); conv = _c(conv); amp = _c(amp)
    amp_out = np.empty_like(delta); d_out = np.empty_like(delta)
    lib().orc_pi_update_linf(_fp(delta), _fp(data), _fp(g), _fp(conv), _fp(amp), ctypes.c_float(alpha), ctypes.c_float(gamma),
                             ctypes.c_float(eps), ctypes.c_float(lo), ctypes.c_float(hi), _fp(amp_out), _fp(d_out),
                             ctypes.c_int64(delta.size))
    return amp_out, d_out


def gra_update(M, last, cur, eta, alpha, delta, data, eps, lo=0.0, hi=1.0):
    """gradient/gra.py:74-93 + :149 — returns (M', delta')"""
    M = _c(M); last = _c(last); cur = _c(cur); delta = _c(delta); data = _c(data)
    M_out = np.empty_like(M); d_out = np.empty_like(M)
    lib().orc_gra_update(_fp(M), _fp(last), _fp(cur), ctypes.c_float(eta), ctypes.c_float(alpha), _fp(delta), _fp(data),
                         ctypes.c_float(eps), ctypes.c_float(lo), ctypes.c_float(hi), _fp(M_out), _fp(d_out), ctypes.c_int64(M.size))
    return M_out, d_out


def adaea_drf(grads, threshold, grad=None):
    """ensemble/adaea.py:115-136, 74-76, 82 — returns (map [B,1,H,W], grad * mask or None)"""
    grads = [_c(g) for g in grads]; grad = _c(grad)
    B, C = grads[0].shape[0], grads[0].shape[1]
    plane = grads[0].size // (B * C)
    arr = (_F * len(grads))(*[_fp(g) for g in grads])
    mp = np.empty((B, 1) + grads[0].shape[2:], np.float32)
    out = np.empty_like(grad) if grad is not None else None
    lib().orc_adaea_drf(arr, len(grads), ctypes.c_float(threshold), _fp(grad), _fp(out), _fp(mp), B, C, ctypes.c_int64(plane))
    return mp, out


def dct_matrices(N):
    """float64 (D, E): D[k][n] = 2 cos(pi (2n+1) k / 2N) — input_transformation/ssm.py:101-133 `dct` with norm=None written as a
    matrix (X = D x) — and E = D^-1 — ssm.py:135-172 `idct`. Pinned against the reference's FFT formulation in
    tests/test_reference_live.py."""
    k = np.arange(N, dtype=np.float64)[:, None]; n = np.arange(N, dtype=np.float64)[None, :]
    D = 2.0 * np.cos(np.pi * (2.0 * n + 1.0) * k / (2.0 * N))
    return D, np.linalg.inv(D)


def spectrum_transform(x, gauss=None, mask=None):
    """ssm.py:41-55 in float64: idct_2d(dct_2d(x + gauss) * mask) = E ((D X D^T) . M) E^T per plane; returns float64"""
    x = np.asarray(x, np.float64)
    N = x.shape[-1]
    D, E = dct_matrices(N)
    X = x if gauss is None else (np.asarray(x, np.float32) + np.asarray(gauss, np.float32)).astype(np.float64)
    Y = D @ X @ D.T
    if mask is not None:
        Y = Y * np.asarray(mask, np.float64)
    return E @ Y @ E.T
