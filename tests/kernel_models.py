"""TEST INFRASTRUCTURE: executable numpy models of the index logic of the second-generation CUDA kernels — the same band
geometry, descriptor tables, accumulator rotation and retire logic as the device code, one Python statement per device step —
so that what can go wrong in those kernels short of arithmetic (a slot reused too early, a warm-up row stored into the previous
band, a destination row written twice or never, a band boundary off by one) is exercised on a box without a GPU against the C
oracle. fp32 fused multiply-adds are emulated through float64 (the 24x24-bit product is exact; one rounding).

  rs_conv_model          transferattack_b200/csrc/dwconv.cu   dwconv_sep_rs_kernel / dwconv_sep_rg_kernel
  rg2_conv_model         … dwconv_sep_rg2_kernel / dwconv_sep_rg3_kernel (paired row weights, tap-exact column pass, zero-row reads)
  colsum_adjoint_model   transferattack_b200/csrc/aten_mean.cu   normalize_bwd_colsum_kernel<FINISH> (per-block trees + ticket)
  dim_tables             transferattack_b200/csrc/dim_direct.cu   host_taps / host_inverse / band table
  dim_fwd_model          … dim_fwd_direct_kernel
  dim_bwd_scatter_model  … dim_bwd_direct_kernel (gather_scatter)
  dim_bwd_gather_model   … dim_bwd_gather_kernel
  dim_bwd_sep_model      … dim_bwd_sep_kernel (the four passes transposed; default adjoint)
"""
import numpy as np

f32 = np.float32
RB = 16


def fma(a, b, c):
    return (np.float64(a) * np.float64(b) + np.float64(c)).astype(np.float32)


# ---- TIM: register-sliding separable convolution -------------------------------------------------------------------------------
def rs_conv_model(g, kcol, krow, bhr=32):
    B, C, H, W = g.shape
    ks = kcol.shape[1]; R = ks // 2; padx = (R + 3) & ~3; off = padx - R; nv = (off + ks + 3 + 3) // 4
    rows = bhr + ks - 1
    out = np.full_like(g, np.nan)
    written = np.zeros(g.shape, np.int32)
    T = W // 4
    for plane in range(B * C):
        c = plane % C
        gp = g.reshape(B * C, H, W)[plane]; op = out.reshape(B * C, H, W)[plane]; wr = written.reshape(B * C, H, W)[plane]
        for band in range((H + bhr - 1) // bhr):
            y0 = band * bhr
            acc = np.zeros((ks, T, 4), f32)
            yl = -(ks - 1); ylim = min(bhr, H - y0)
            for r in range(rows):
                yy = y0 - R + r
                win = np.zeros((T, 4 * nv), f32)                       # predicated 128-bit loads: zero outside the image
                if 0 <= yy < H:
                    for t in range(T):
                        for k in range(nv):
                            col = 4 * t - padx + 4 * k
                            if 0 <= col < W:
                                win[t, 4 * k:4 * k + 4] = gp[yy, col:col + 4]
                t4 = np.zeros((T, 4), f32)
                for j in range(ks):
                    for cc in range(4):
                        t4[:, cc] = fma(krow[c, j], win[:, off + j + cc], t4[:, cc])
                rr = r % ks                                            # position inside the ks-fold unrolled group
                for i in range(ks):
                    s = ((rr - i) % ks + ks) % ks
                    acc[s] = fma(kcol[c, i], t4, acc[s])
                sc = (rr + 1) % ks
                if 0 <= yl < ylim:
                    op[y0 + yl] = acc[sc].reshape(-1); wr[y0 + yl] += 1
                acc[sc] = 0
                yl += 1
    assert (written == 1).all(), "an output row was stored twice or never"
    return out


def rg2_conv_model(g, kcol, krow, bhr=32):
    """dwconv_sep_rg2_kernel (round 2): the band's bhr + ks - 1 rows fully unrolled; per input row
      * row pass with the DATA broadcast and the WEIGHTS paired: pair (out[x], out[x+1]) += (w[m], w[m-1]) * v[OFF + x + m], m = 1 .. ks-1,
        the two end taps (m = 0 on the low lane, m = ks on the high lane) as scalar fmas;
      * column pass only for the taps that exist: input row r is tap i of output row y = r - i, 0 <= y < bhr; tap 0 starts from +0;
      * rows above / below the image are skipped (only possible in the first / last band: H % bhr == 0); a slot whose first tap
        was skipped is set to +0 instead; window quarters outside the row read zeros (the device reads its static zero rows);
      * output row y leaves after input row y + ks - 1.  Shared (channel-independent) factors, W % 4 == 0, H % bhr == 0."""
    B, C, H, W = g.shape
    ks = kcol.shape[1]; R = ks // 2; padx = (R + 3) & ~3; off = padx - R; nv = (off + ks + 3 + 3) // 4
    assert H % bhr == 0 and W % 4 == 0
    rows = bhr + ks - 1
    nbands = H // bhr
    out = np.full_like(g, np.nan)
    written = np.zeros(g.shape, np.int32)
    T = W // 4
    kr, kc = krow[0], kcol[0]
    wp = [(kr[m] if m < ks else f32(0), kr[m - 1] if m >= 1 else f32(0)) for m in range(ks + 1)]
    for plane in range(B * C):
        gp = g.reshape(B * C, H, W)[plane]; op = out.reshape(B * C, H, W)[plane]; wr = written.reshape(B * C, H, W)[plane]
        for band in range(nbands):
            y0 = band * bhr
            top_ok, bot_ok = band > 0, band < nbands - 1
            acc = np.full((ks, T, 4), np.nan, f32)                     # no slot may be read before its first tap (or its +0) was written
            for r in range(rows):
                rv = top_ok if r < R else (bot_ok if r >= rows - R else True)
                if rv:
                    yy = y0 - R + r
                    assert 0 <= yy < H
                    win = np.zeros((T, 4 * nv), f32)
                    for t in range(T):
                        for k in range(nv):
                            col = 4 * t - padx + 4 * k
                            if 0 <= col < W:
                                win[t, 4 * k:4 * k + 4] = gp[yy, col:col + 4]
                    t4 = np.zeros((T, 4), f32)
                    for pair in (0, 2):                                # pair 0 = outputs (0, 1), pair 1 = outputs (2, 3)
                        lo = fma(kr[0], win[:, off + pair], f32(0)); hi = np.zeros(T, f32)
                        for m in range(1, ks):
                            v = win[:, off + pair + m]
                            lo = fma(wp[m][0], v, lo); hi = fma(wp[m][1], v, hi)
                        hi = fma(kr[ks - 1], win[:, off + pair + ks], hi)
                        t4[:, pair] = lo; t4[:, pair + 1] = hi
                    for i in range(ks):
                        y = r - i
                        if 0 <= y < bhr:
                            s = y % ks
                            acc[s] = fma(kc[i], t4, np.zeros_like(t4) if i == 0 else acc[s])
                elif r < bhr:
                    acc[r % ks] = 0
                if r >= ks - 1:
                    y = r - (ks - 1)
                    assert not np.isnan(acc[y % ks]).any()
                    op[y0 + y] = acc[y % ks].reshape(-1); wr[y0 + y] += 1
    assert (written == 1).all(), "an output row was stored twice or never"
    return out


def colsum_adjoint_model(gout, std, cfg):
    """normalize_bwd_colsum_kernel<FINISH = true>: CTA (x, b) = ATen's virtual block x of sample b; thread t owns the 128-bit vectors
    col, col + S, ... (col = 512 x + t), divides by std[channel], accumulates |.| component-wise; the CTA runs block_x_reduce /
    block_y_reduce on its 512 values; the per-block partials are added by global_reduce's final tree. Returns (gin, mean)."""
    B = gout.shape[0]; C = gout.shape[1]
    n = gout[0].size
    bw, bh, cpo = cfg["bw"], cfg["bh"], cfg["cpo"]; S = bw * bh * cpo
    assert bw * bh == 512
    nvec = n // 4; plane_vec = nvec // C
    K = bw // 32
    gin = np.empty_like(gout).reshape(B, nvec, 4)
    src = gout.reshape(B, nvec, 4)
    mean = np.zeros(B, f32)
    for b in range(B):
        partial = np.zeros(cpo, f32)
        for x in range(cpo):
            vals = np.zeros(512, f32)
            for t in range(512):
                col = 512 * x + t
                a = np.zeros(4, f32)
                v = col
                while v < nvec:
                    ch = int(v >= plane_vec) + int(v >= 2 * plane_vec) + int(v >= 3 * plane_vec)
                    q = (src[b, v] / std[ch]).astype(f32)
                    gin[b, v] = q
                    a = (a + np.abs(q)).astype(f32)
                    v += S
                vals[t] = f32(f32(f32(a[0] + a[1]) + a[2]) + a[3])
            rows = np.zeros(bh, f32)
            for ty in range(bh):                                        # block_x_reduce of row ty: lane l holds tx = l + 32 k
                a = [vals[ty * bw + np.arange(32) + 32 * k].copy() for k in range(K)]
                rows[ty] = _x_tree(a, K)
            h = bh // 2
            while h >= 1:                                               # block_y_reduce
                rows[:h] = (rows[:h] + rows[h:2 * h]).astype(f32); h //= 2
            partial[x] = rows[0]
        lanes = np.zeros(32 * max(K, 1), f32); lanes[:cpo] = partial     # global_reduce: partial i at tx = i, ty = 0
        a = [lanes[np.arange(32) + 32 * k].copy() for k in range(K)]
        tot = _x_tree(a, K)
        mean[b] = f32(tot * f32(f32(B) / f32(B * n)))
    return gin.reshape(gout.shape), mean


# ---- DIM: host tables ------------------------------------------------------------------------------------------------------------
def host_taps(inn, out):
    scale = f32(inn) / f32(out)
    t = []
    for d in range(out):
        src = fma(scale, f32(d) + f32(0.5), f32(-0.5))
        if src < 0:
            src = f32(0)
        i0 = int(src); i1 = i0 + (1 if i0 < inn - 1 else 0)
        t.append((i0, i1, f32(src - f32(i0))))
    return t


def host_inverse(t, out, inn):
    inv = [[0, 0] for _ in range(inn)]
    for d in range(out):
        for idx in (t[d][0], t[d][1]):
            e = inv[idx]
            if e[1] == 0:
                e[0] = d; e[1] = 1
            else:
                e[1] = d - e[0] + 1
    return inv


def dim_tables(S, rnd, R):
    t2, t1 = host_taps(R, S), host_taps(S, rnd)
    return t2, t1, host_inverse(t2, S, R), host_inverse(t1, rnd, S)


def _hl(w0, w1, a, b):          # blend mode 1: fma(w0, a, w1*b)
    return fma(w0, a, f32(w1) * f32(b))


def tap_w(t, i):
    w = f32(0)
    if t[0] == i:
        w = f32(1) - t[2]
    if t[1] == i:
        w = f32(w + t[2])
    return w


def dim_fwd_model(x, rnd, R, top, left):
    P, S, _ = x.shape
    t2, t1, _, _ = dim_tables(S, rnd, R)
    out = np.full_like(x, np.nan)
    for pl in range(P):
        for oy0 in range(0, S, RB):
            oy1 = min(oy0 + RB, S) - 1; nb = oy1 - oy0 + 1
            pr0, pr1 = t2[oy0][0], t2[oy1][1]
            q0, q1 = max(pr0 - top, 0), min(pr1 - top, rnd - 1)
            nq = q1 - q0 + 1 if q0 <= q1 else 0
            bufC = np.full((nq + 1, rnd + 1), np.nan, f32); bufC[nq, :] = 0; bufC[:nq, rnd] = 0      # zero row / zero column
            if nq:
                sr0 = t1[q0][0]; nsr = t1[q1][1] - sr0 + 1; bufA = x[pl, sr0:sr0 + nsr]
                for col in range(rnd):
                    i0, i1, wl1 = t1[col]; wl0 = f32(1) - wl1
                    for q in range(nq):
                        a0, a1, l1 = t1[q0 + q]; l0 = f32(1) - l1
                        t = _hl(wl0, wl1, bufA[a0 - sr0, i0], bufA[a0 - sr0, i1]); b = _hl(wl0, wl1, bufA[a1 - sr0, i0], bufA[a1 - sr0, i1])
                        bufC[q, col] = _hl(l0, l1, t, b)
            for col in range(S):
                i0, i1, wl1 = t2[col]; wl0 = f32(1) - wl1
                xa, xb = i0 - left, i1 - left
                xa = xa if 0 <= xa < rnd else rnd; xb = xb if 0 <= xb < rnd else rnd
                for r in range(nb):
                    a0, a1, l1 = t2[oy0 + r]; l0 = f32(1) - l1
                    ya, yb = a0 - top - q0, a1 - top - q0
                    ya = ya if 0 <= ya < nq else nq; yb = yb if 0 <= yb < nq else nq
                    t = _hl(wl0, wl1, bufC[ya, xa], bufC[ya, xb]); b = _hl(wl0, wl1, bufC[yb, xa], bufC[yb, xb])
                    out[pl, oy0 + r, col] = _hl(l0, l1, t, b)
    return out


def _band(sy0, S, top, inv1, inv2):
    sy1 = min(sy0 + RB, S) - 1
    q0, q1 = 10 ** 9, -1
    for sy in range(sy0, sy1 + 1):
        lo, c = inv1[sy]
        if c > 0:
            q0 = min(q0, lo); q1 = max(q1, lo + c - 1)
    nq = q1 - q0 + 1 if q0 <= q1 else 0
    if not nq:
        q0 = 0
    oyA, oyB = 10 ** 9, -1
    for q in range(nq):
        lo, c = inv2[q0 + q + top]
        if c > 0:
            oyA = min(oyA, lo); oyB = max(oyB, lo + c - 1)
    nu = oyB - oyA + 1 if oyA <= oyB else 0
    if not nu:
        oyA = 0
    return sy1, q0, nq, oyA, nu


def _gather_scatter(src, nrows, desc, tab, lo, cnt, col, emit_lo, emit_hi, emit):
    """gather_scatter of dim_direct.cu: horizontal inverse-range sum per source row, scatter into two rotating accumulators,
    rows retired when the monotone tap index moves past them"""
    pcur = desc[0][0] if nrows > 0 else emit_hi + 1
    p = emit_lo
    while p < pcur and p <= emit_hi:
        emit(p, f32(0)); p += 1
    accA, accB = f32(0), f32(0)
    for r in range(nrows):
        i0, i1, l1 = desc[r]
        h = f32(0)
        for k in range(cnt):
            h = fma(tap_w(tab[lo + k], col), src[r, lo + k], h)
        while pcur < i0:
            if emit_lo <= pcur <= emit_hi:
                emit(pcur, accA)
            accA, accB = accB, f32(0); pcur += 1
        l0 = f32(1) - l1
        accA = fma(l0, h, accA)
        if i1 == i0:
            accA = fma(l1, h, accA)
        else:
            accB = fma(l1, h, accB)
    if nrows > 0:
        if emit_lo <= pcur <= emit_hi:
            emit(pcur, accA)
        if emit_lo <= pcur + 1 <= emit_hi:
            emit(pcur + 1, accB)
        for p in range(max(pcur + 2, emit_lo), emit_hi + 1):
            emit(p, f32(0))


def dim_bwd_scatter_model(g, rnd, R, top, left):
    P, S, _ = g.shape
    t2, t1, inv2, inv1 = dim_tables(S, rnd, R)
    gin = np.full_like(g, np.nan); hits = np.zeros(g.shape, np.int32)
    for pl in range(P):
        for sy0 in range(0, S, RB):
            sy1, q0, nq, oyA, nu = _band(sy0, S, top, inv1, inv2)
            bufU = g[pl, oyA:oyA + nu]
            bufG = np.full((max(nq, 1), rnd), np.nan, f32); wr = np.zeros((max(nq, 1), rnd), np.int32)
            if nq:
                for col in range(rnd):
                    px = col + left; lo, c = inv2[px]

                    def emit(p, v, col=col):
                        bufG[p - top - q0, col] = v; wr[p - top - q0, col] += 1
                    _gather_scatter(bufU, nu, [t2[oyA + r] for r in range(nu)], t2, lo, c, px, q0 + top, q0 + nq - 1 + top, emit)
                assert (wr[:nq] == 1).all(), "a g1 row was written twice or never"
            for col in range(S):
                lo, c = inv1[col]

                def emit2(s, v, col=col):
                    gin[pl, s, col] = v; hits[pl, s, col] += 1
                _gather_scatter(bufG, nq, [t1[q0 + q] for q in range(nq)], t1, lo, c, col, sy0, sy1, emit2)
    assert (hits == 1).all(), "a gin element was written twice or never"
    return gin


def dim_bwd_gather_model(g, rnd, R, top, left):
    P, S, _ = g.shape
    t2, t1, inv2, inv1 = dim_tables(S, rnd, R)
    gin = np.full_like(g, np.nan)

    def gather(src, rowlo, roww, collo, colw):
        acc = f32(0)
        for a, wy in enumerate(roww):
            h = f32(0)
            for b, wx in enumerate(colw):
                h = fma(wx, src[rowlo + a, collo + b], h)
            acc = fma(wy, h, acc)
        return acc
    for pl in range(P):
        for sy0 in range(0, S, RB):
            sy1, q0, nq, oyA, nu = _band(sy0, S, top, inv1, inv2)
            bufU = g[pl, oyA:oyA + nu]
            bufG = np.full((max(nq, 1), rnd), np.nan, f32)
            for q in range(nq):
                p = q0 + q + top; lo, c = inv2[p]; roww = [tap_w(t2[lo + a], p) for a in range(c)]
                for col in range(rnd):
                    px = col + left; lx, cx = inv2[px]
                    bufG[q, col] = gather(bufU, lo - oyA, roww, lx, [tap_w(t2[lx + b], px) for b in range(cx)])
            for r in range(sy1 - sy0 + 1):
                sy = sy0 + r; lo, c = inv1[sy]; roww = [tap_w(t1[lo + a], sy) for a in range(c)]
                for col in range(S):
                    lx, cx = inv1[col]
                    gin[pl, sy, col] = gather(bufG, lo - q0, roww, lx, [tap_w(t1[lx + b], col) for b in range(cx)]) if nq else f32(0)
    return gin


def dim_bwd_sep_model(g, rnd, R, top, left):
    """dim_bwd_sep_kernel (the default adjoint in round 2): per band of RB source rows the four passes transposed, every one a
    gather over an inverse range with the nan-poisoned buffers the device uses (a read outside what a pass wrote shows up as NaN):
      vT2  V[q][ox]  = sum_a w(oy = lo + a, p = q0 + q + top) * U[lo + a - oyA][ox]       (rows: the band's y1 rows as y2 rows)
      hT2  G[q][qx]  = sum_b w(ox = lo + b, px = qx + left)   * V[q][lo + b]               (crop = the pad's adjoint)
      vT1  W[r][qx]  = sum_a w(q = lo + a, sy = sy0 + r)      * G[lo + a - q0][qx]
      hT1  gin[sy][sx] = sum_b w(qx = lo + b, sx)             * W[r][lo + b]"""
    P, S, _ = g.shape
    t2, t1, inv2, inv1 = dim_tables(S, rnd, R)
    gin = np.full_like(g, np.nan)
    wrote = np.zeros(g.shape, np.int32)
    for pl in range(P):
        for sy0 in range(0, S, RB):
            sy1, q0, nq, oyA, nu = _band(sy0, S, top, inv1, inv2)
            nb = sy1 - sy0 + 1
            if nq == 0:
                gin[pl, sy0:sy1 + 1] = 0; wrote[pl, sy0:sy1 + 1] += 1
                continue
            U = g[pl, oyA:oyA + nu]
            V = np.full((nq, S), np.nan, f32)
            for q in range(nq):
                p = q0 + q + top; lo, c = inv2[p]
                acc = np.zeros(S, f32)
                for a in range(c):
                    assert 0 <= lo + a - oyA < nu
                    acc = fma(tap_w(t2[lo + a], p), U[lo + a - oyA], acc)
                V[q] = acc
            G = np.full((nq, rnd), np.nan, f32)
            for qx in range(rnd):
                px = qx + left; lo, c = inv2[px]
                for q in range(nq):
                    h = f32(0)
                    for b in range(c):
                        h = fma(tap_w(t2[lo + b], px), V[q, lo + b], h)
                    G[q, qx] = h
            Wb = np.full((nb, rnd), np.nan, f32)
            for r in range(nb):
                sy = sy0 + r; lo, c = inv1[sy]
                acc = np.zeros(rnd, f32)
                for a in range(c):
                    assert 0 <= lo + a - q0 < nq
                    acc = fma(tap_w(t1[lo + a], sy), G[lo + a - q0], acc)
                Wb[r] = acc
            for sx in range(S):
                lo, c = inv1[sx]
                for r in range(nb):
                    h = f32(0)
                    for b in range(c):
                        h = fma(tap_w(t1[lo + b], sx), Wb[r, lo + b], h)
                    gin[pl, sy0 + r, sx] = h; wrote[pl, sy0 + r, sx] += 1
    assert (wrote == 1).all(), "a source element was written twice or never"
    return gin


# ---- csrc/aten_mean.cuh: the TA_MEAN_TORCH replay, as the kernel executes it --------------------------------------------
def _shfl_down(v, off, width=32):
    """__shfl_down_sync(full, v, off) for one warp: lane l reads lane l + off, or itself when that is past the warp"""
    out = v.copy()
    for l in range(32):
        if (l % width) + off < width:
            out[l] = v[l + off]
    return out


def _x_tree(a, K):
    """aten_x_tree<16>: a[k][lane] = value[tx = lane + 32k]; register halving (the shared-memory levels), then shfl_down 16..1"""
    f32 = np.float32
    h = 8
    while h >= 1:
        if h < K:
            for k in range(h):
                a[k] = (a[k] + a[k + h]).astype(f32)
        h >>= 1
    v = a[0]
    for o in (16, 8, 4, 2, 1):
        v = (v + _shfl_down(v, o)).astype(f32)
    return v[0]


def aten_mean_kernel_model(x, cfg, cl=8):
    """mean-sum of one sample x [n] (already |g|) exactly as csrc/aten_mean.cuh + the fused cluster kernel compute it: rows of S
    128-bit vectors, one thread per vector column (the 4 accumulators are the 4 components, rows in order), CTA r owning
    columns [r*W4, (r+1)*W4); then aten_tree_mean: a warp per (virtual block, ty) row — lanes hold tx = lane + 32k —, one
    thread per virtual block for the y tree, the final tree over the cpo block sums. `cl` only changes who owns which column
    (the DSMEM gather reads virtual thread vt's value from CTA vt // W4)."""
    f32 = np.float32
    n = x.shape[0]
    bw, bh, cpo, S = cfg["bw"], cfg["bh"], cfg["cpo"], cfg["stride"]
    assert n % 4 == 0 and S % cl == 0
    nvec = n // 4
    X = x.reshape(nvec, 4)
    val = np.zeros(S, f32)
    for t in range(S):                                    # phase 1 (the owner CTA is t // (S // cl); irrelevant for the value)
        a = np.zeros(4, f32)
        for v in range(t, nvec, S):
            a = (a + X[v]).astype(f32)
        val[t] = f32(f32(f32(a[0] + a[1]) + a[2]) + a[3])
    K = bw >> 5
    lanes = np.arange(32)
    s_row = np.zeros(cpo * bh, f32)
    for row in range(cpo * bh):                           # phase 2a: x tree per (virtual block, ty) row
        base = row * bw
        a = [np.zeros(32, f32) for _ in range(16)]
        for k in range(K):
            a[k] = val[base + lanes + 32 * k].astype(f32)
        s_row[row] = _x_tree(a, K)
    s_blk = np.zeros(max(cpo, 1), f32)
    for cb in range(cpo):                                 # phase 2b: y tree, one thread per virtual block
        a = [s_row[cb * bh + y] if y < bh else f32(0) for y in range(16)]
        h = 8
        while h >= 1:
            if h < bh:
                for y in range(h):
                    a[y] = f32(a[y] + a[y + h])
            h >>= 1
        s_blk[cb] = a[0]
    if cpo == 1:
        return s_blk[0]
    a = [np.zeros(32, f32) for _ in range(16)]            # phase 2c: global_reduce's last block
    for k in range(K):
        idx = lanes + 32 * k
        a[k] = np.where(idx < cpo, s_blk[np.minimum(idx, cpo - 1)], f32(0)).astype(f32)
    return _x_tree(a, K)
