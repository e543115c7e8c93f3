"""TESTS ONLY: a stand-in for ``transferattack_b200.ops.CudaBackend`` that serves the same method surface from the
C oracle on CPU tensors. It lets the host-side logic of the package (loop control flow, hook argument tolerance, RNG
consumption order, plugin classes, multi-process sharding) run on a box without a GPU and be compared bit for bit
with the reference. The product never installs it (``ops._install_backend_for_tests`` is only called from tests/)."""
import numpy as np
import torch

import oracle


def _np(t):
    return None if t is None else t.detach().cpu().contiguous().numpy()


def _t(a, like=None):
    return torch.from_numpy(np.ascontiguousarray(a))


class OracleBackend:
    calls = None

    def __init__(self):
        self.calls = []

    def _log(self, name):
        self.calls.append(name)

    def abs_mean(self, g, mode=0):
        self._log("abs_mean")
        if mode == 1:
            from oracle import aten_reduce
            gn = _np(g)
            r = aten_reduce.emulate_numpy(np.abs(gn).reshape(gn.shape[0], -1))
            return None if r is None else _t(r)
        return _t(oracle.abs_mean_per_sample(_np(g)))

    def momentum(self, g, m, scale, decay, out=None):
        self._log("momentum")
        r = _t(oracle.momentum(_np(g), _np(m), _np(scale).reshape(-1), float(decay)))
        if out is not None:
            out.copy_(r); return out
        return r

    def update_linf(self, delta, data, direction, alpha, eps, lo, hi, alpha_t=None, dir_mode=0, out=None):
        self._log("update_linf")
        r = _t(oracle.update_linf(_np(delta), _np(data), _np(direction), float(alpha), float(eps), float(lo), float(hi),
                                  alpha_t=_np(alpha_t), dir_mode=dir_mode))
        if out is not None:
            out.copy_(r); return out
        return r

    def update_l2(self, delta, data, g, alpha, eps, lo, hi):
        self._log("update_l2")
        return _t(oracle.update_l2(_np(delta), _np(data), _np(g), float(alpha), float(eps), float(lo), float(hi)))

    def clamp_box(self, delta, data, lo, hi):
        self._log("clamp_box")
        return _t(oracle.clamp_box(_np(delta), _np(data), float(lo), float(hi)))

    def init_l2_scale(self, delta, r, data, eps, lo, hi):
        self._log("init_l2_scale")
        return _t(oracle.init_l2_scale(_np(delta), _np(r), _np(data), float(eps), float(lo), float(hi)))

    def fused_tail(self, g, m, m_out, delta, delta_out, data, xadv_out, scale, scale_out, decay, alpha, eps, lo, hi,
                   mean_mode=0, addend=None, gbar_out=None, mean=None, std=None, emit_normalized=False, grad_wrt_xn=False):
        """ta_fused_tail composed from the oracle's single ops, in the kernel's order: g / std, + addend, mean, momentum,
        update, x + delta', Normalize."""
        self._log("fused_tail_nf" if emit_normalized else "fused_tail")
        gn = _np(g)
        if grad_wrt_xn:
            gn = oracle.normalize_bwd(gn, np.asarray(std, np.float32))
        if addend is not None:
            gn = oracle.add(gn, _np(addend))
        if scale is not None:
            sc = _np(scale).reshape(-1)
        elif mean_mode == 1:        # TA_MEAN_TORCH: the summation tree of torch's CUDA kernel on a 148-SM device
            from oracle import aten_reduce
            sc = aten_reduce.emulate_numpy(np.abs(gn).reshape(gn.shape[0], -1))
            if sc is None:
                return False
        else:
            sc = oracle.abs_mean_per_sample(gn)
        mo, do, xo = oracle.fused_update_linf(gn, _np(m), _np(delta), _np(data), sc, float(decay), float(alpha), float(eps),
                                              float(lo), float(hi), want_xadv=xadv_out is not None)
        if emit_normalized:
            xo = oracle.normalize_fwd(xo, np.asarray(mean, np.float32), np.asarray(std, np.float32))
        with torch.no_grad():
            m_out.copy_(_t(mo)); delta_out.copy_(_t(do))
            if xadv_out is not None:
                xadv_out.copy_(_t(xo))
            if scale_out is not None:
                scale_out.copy_(_t(sc))
            if gbar_out is not None:
                B = gn.shape[0]
                gbar_out.copy_(_t((gn.reshape(B, -1) / sc.reshape(B, 1).astype(np.float32)).astype(np.float32).reshape(gn.shape)))
        return True

    def fused_update_linf(self, g, m, m_out, delta, delta_out, data, xadv_out, scale, scale_out, decay, alpha, eps, lo, hi,
                          mean_mode=0):
        self.fused_tail(g, m, m_out, delta, delta_out, data, xadv_out, scale, scale_out, decay, alpha, eps, lo, hi, mean_mode)

    def fused_update_linf_nf(self, g, m, m_out, delta, delta_out, data, xn_out, scale, scale_out, decay, alpha, eps, lo, hi,
                             mean, std, grad_wrt_xn, mean_mode=0):
        return self.fused_tail(g, m, m_out, delta, delta_out, data, xn_out, scale, scale_out, decay, alpha, eps, lo, hi, mean_mode,
                               mean=mean, std=std, emit_normalized=True, grad_wrt_xn=grad_wrt_xn)

    def stage_add(self, data, delta, look=None, coef=0.0, out=None):
        self._log("stage_add")
        d = _np(data)
        if delta is None:   # x + coef*look on an already formed x
            r = d if look is None else (d + (np.float32(coef) * _np(look)).astype(np.float32)).astype(np.float32)
            return _t(r)
        return _t(oracle.stage_add(d, _np(delta), _np(look), float(coef)))

    def neighbor_stage(self, data, delta, noise, look=None, coef=0.0, out=None):
        self._log("neighbor_stage")
        return _t(oracle.neighbor_stage(_np(data), _np(delta), _np(noise), _np(look), float(coef)))

    def normalize(self, x, mean, std, forward=True):
        self._log("normalize")
        if forward:
            return _t(oracle.normalize_fwd(_np(x), _np(mean), _np(std)))
        return _t(oracle.normalize_bwd(_np(x), _np(std)))

    def sim(self, x, S, forward=True):
        self._log("sim")
        return _t(oracle.sim_fwd(_np(x), S) if forward else oracle.sim_bwd(_np(x), S))

    def admix(self, x, perm, strength, S, A, forward=True):
        self._log("admix")
        if forward:
            return _t(oracle.admix_fwd(_np(x), _np(perm), float(strength), S))
        return _t(oracle.admix_bwd(_np(x), S, A))

    def dim(self, x, rnd, R, top, left, forward=True):
        self._log("dim")
        if forward:
            return _t(oracle.dim_fwd(_np(x), int(rnd), int(R), int(top), int(left), blend=1))
        return _t(oracle.dim_bwd(_np(x), int(rnd), int(R), int(top), int(left)))

    def dwconv2d(self, g, k):
        self._log("dwconv2d")
        return _t(oracle.dwconv2d(_np(g), _np(k)))

    def dwconv2d_sep(self, g, kcol, krow, host=None):
        self._log("dwconv2d_sep")
        return _t(oracle.dwconv2d_sep(_np(g), _np(kcol), _np(krow)))

    def pi_cut_noise(self, amp, momentum, coef, eps):
        self._log("pi_cut_noise")
        a, c = oracle.pi_cut_noise(_np(amp), _np(momentum), float(coef), float(eps))
        return _t(a), _t(c)

    def pi_update_linf(self, delta, data, g, conv, amp, alpha, gamma, eps, lo, hi):
        self._log("pi_update_linf")
        a, d = oracle.pi_update_linf(_np(delta), _np(data), _np(g), _np(conv), _np(amp), float(alpha), float(gamma), float(eps),
                                     float(lo), float(hi))
        return _t(a), _t(d)

    def gra_update(self, M, last, cur, eta, alpha, delta, data, eps, lo, hi):
        self._log("gra_update")
        m, d = oracle.gra_update(_np(M), _np(last), _np(cur), float(eta), float(alpha), _np(delta), _np(data), float(eps), float(lo), float(hi))
        return _t(m), _t(d)

    def adaea_drf(self, grads, threshold, grad=None, want_map=False):
        self._log("adaea_drf")
        mp, out = oracle.adaea_drf([_np(g) for g in grads], float(threshold), _np(grad))
        return (None if out is None else _t(out)), (_t(mp) if (want_map or grad is None) else None)

    def spectrum_transform(self, x, gauss, mask, precision=1):
        self._log("spectrum_transform")
        return _t(oracle.spectrum_transform(_np(x), _np(gauss), _np(mask)).astype(np.float32))

    def lin_sample(self, x, gbar, coefs, forward=True):
        self._log("lin_sample")
        if forward:
            return _t(oracle.lin_sample_fwd(_np(x), _np(gbar), np.asarray(coefs, np.float32)))
        return _t(oracle.lin_sample_bwd(_np(x), len(coefs)))

    def accumulate(self, acc, g, first):
        self._log("accumulate")
        return _t(oracle.accumulate(_np(acc), _np(g), first))

    def variance_finalize(self, acc, cur, num_neighbor):
        self._log("variance_finalize")
        return _t(oracle.variance_finalize(_np(acc), _np(cur), num_neighbor))

    def add(self, a, b):
        self._log("add")
        return _t(oracle.add(_np(a), _np(b)))

    def quantize_u8(self, data, delta, to_nhwc=True):
        self._log("quantize_u8")
        return torch.from_numpy(oracle.quantize_u8(_np(data), _np(delta), to_nhwc))
