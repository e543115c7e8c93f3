/*
 * ta_oracle.c — TEST INFRASTRUCTURE ONLY. CPU restatement (plain C, IEEE fp32, one rounding per
 * reference op) of the per-iteration arithmetic of TransferAttack's hot loop. It is the checker the
 * CUDA kernels of libta_b200.so are compared against; nothing in transferattack_b200/ may import,
 * link or call it (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs).
 *
 * Parity pin: the reference has no tests or golden vectors of its own, so this file
 * is pinned against outputs of the reference itself: tests/golden/make_golden.py imports the
 * unmodified reference classes from /root/reference, runs them on seeded inputs on the CPU and stores
 * inputs + outputs under tests/golden/*.npz; tests/test_oracle_golden.py checks every function below
 * against those files (bit-exact where the reference op order is fully determined, else within the
 * stated tolerance).
 *
 * Build:  gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fPIC -shared ta_oracle.c -lm
 *         (-ffp-contract=off: the compiler must not fuse a*b+c; the only FMA is the explicit fmaf
 *          in the bilinear source index, which is what ATen computes.)
 *
 * Citations are file:line under the reference's transferattack/ directory.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* torch.sign: (0 < x) - (x < 0); sign(NaN) = 0, sign(+-0) = 0 */
static inline float sgnf(float v) { return (float)((0.0f < v) - (v < 0.0f)); }
/* torch.max / torch.min / torch.clamp propagate NaN from either operand */
static inline float max_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
static inline float min_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

/* ---- attack.py:128  grad.abs().mean(dim=(1,2,3)) — TA_MEAN_EXACT definition ------------------ */
ORC_API void orc_abs_mean_per_sample(const float* g, float* mean_out, int B, int64_t n) {
  for (int b = 0; b < B; ++b) {
    double s = 0.0;
    const float* p = g + (int64_t)b * n;
    for (int64_t i = 0; i < n; ++i) s += (double)fabsf(p[i]);
    mean_out[b] = (float)(s / (double)n);
  }
}

/* ---- attack.py:128  momentum * decay + grad / mean  (scale given: [B]) -------------------------- */
ORC_API void orc_momentum(const float* g, const float* m, const float* scale, float decay, float* m_out,
                          int B, int64_t n) {
  for (int b = 0; b < B; ++b) {
    const float mu = scale[b];
    for (int64_t i = 0; i < n; ++i) {
      const int64_t j = (int64_t)b * n + i;
      const float t1 = m ? m[j] * decay : 0.0f; /* python `0 * decay` = 0 */
      const float t2 = g[j] / mu;
      m_out[j] = t1 + t2;
    }
  }
}

/* ---- attack.py:147,152  L-inf update + utils.py:68-69 clamp ------------------------------------- */
ORC_API void orc_update_linf(const float* delta, const float* data, const float* dir, const float* alpha_t,
                             float alpha, float eps, float lo, float hi, int dir_mode, float* delta_out,
                             int64_t N) {
  const float neg_eps = -eps;
  for (int64_t j = 0; j < N; ++j) {
    const float d = (dir_mode == 0) ? sgnf(dir[j]) : dir[j];
    const float a = alpha_t ? alpha_t[j] : alpha;
    const float st = a * d;
    const float d1 = delta[j] + st;
    const float d2 = min_nan(max_nan(d1, neg_eps), eps); /* torch.clamp(x, -eps, eps) */
    const float l = lo - data[j];
    const float h = hi - data[j];
    delta_out[j] = min_nan(max_nan(d2, l), h);
  }
}

/* ---- attack.py:141 ------------------------------------------------------------------------------- */
ORC_API void orc_clamp_box(const float* delta, const float* data, float lo, float hi, float* out, int64_t N) {
  for (int64_t j = 0; j < N; ++j) out[j] = min_nan(max_nan(delta[j], lo - data[j]), hi - data[j]);
}

/* ---- attack.py:148-152  L2 update. Norms accumulate in fp64 (torch's order is not reproducible;
 *      compared with tolerance). renorm: rows with norm > maxnorm scaled by maxnorm / (norm + 1e-7). */
ORC_API void orc_update_l2(const float* delta, const float* data, const float* g, float alpha, float eps,
                           float lo, float hi, float* delta_out, int B, int64_t n) {
  for (int b = 0; b < B; ++b) {
    const int64_t o = (int64_t)b * n;
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += (double)g[o + i] * (double)g[o + i];
    const float gn = (float)sqrt(s);
    const float den = gn + 1e-20f;
    double s2 = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      const float gh = g[o + i] / den;
      const float y = delta[o + i] + gh * alpha;
      delta_out[o + i] = y;
      s2 += (double)y * (double)y;
    }
    const float yn = (float)sqrt(s2);
    if (yn > eps) {
      const float f = eps / (yn + 1e-7f);
      for (int64_t i = 0; i < n; ++i) delta_out[o + i] = delta_out[o + i] * f;
    }
    for (int64_t i = 0; i < n; ++i)
      delta_out[o + i] = min_nan(max_nan(delta_out[o + i], lo - data[o + i]), hi - data[o + i]);
  }
}

/* ---- attack.py:136-141  L2 random start: delta *= r / n * eps ; clamp --------------------------- */
ORC_API void orc_init_l2_scale(const float* delta, const float* r, const float* data, float eps, float lo,
                               float hi, float* out, int B, int64_t n) {
  for (int b = 0; b < B; ++b) {
    const int64_t o = (int64_t)b * n;
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += (double)delta[o + i] * (double)delta[o + i];
    const float nn = (float)sqrt(s);
    for (int64_t i = 0; i < n; ++i) {
      const float f = (r[o + i] / nn) * eps;
      const float v = delta[o + i] * f;
      out[o + i] = min_nan(max_nan(v, lo - data[o + i]), hi - data[o + i]);
    }
  }
}

/* ---- the fused iteration tail: attack.py:128 + :147,152 + next :88 ------------------------------- */
ORC_API void orc_fused_update_linf(const float* g, const float* m, float* m_out, const float* delta,
                                   float* delta_out, const float* data, float* xadv_out, const float* scale,
                                   float decay, float alpha, float eps, float lo, float hi, int B, int64_t n) {
  const float neg_eps = -eps;
  for (int b = 0; b < B; ++b) {
    const float mu = scale[b];
    for (int64_t i = 0; i < n; ++i) {
      const int64_t j = (int64_t)b * n + i;
      const float t1 = m ? m[j] * decay : 0.0f;
      const float t2 = g[j] / mu;
      const float mm = t1 + t2;
      const float st = alpha * sgnf(mm);
      const float d1 = delta[j] + st;
      const float d2 = min_nan(max_nan(d1, neg_eps), eps);
      const float x = data[j];
      const float dn = min_nan(max_nan(d2, lo - x), hi - x);
      m_out[j] = mm;
      delta_out[j] = dn;
      if (xadv_out) xadv_out[j] = x + dn;
    }
  }
}

/* ---- attack.py:88 / gradient/nifgsm.py:39 ---------------------------------------------------------- */
ORC_API void orc_stage_add(const float* data, const float* delta, const float* look, float coef, float* out,
                           int64_t N) {
  for (int64_t j = 0; j < N; ++j) {
    const float x = data[j] + delta[j];
    if (look) {
      const float t = coef * look[j];
      out[j] = x + t;
    } else {
      out[j] = x;
    }
  }
}

/* ---- utils.py:72-79  Normalize = clone; sub_(mean); div_(std) ---------------------------------------- */
ORC_API void orc_normalize_fwd(const float* x, const float* mean, const float* std, float* out, int B, int C,
                               int64_t plane) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const int64_t o = ((int64_t)b * C + c) * plane;
      for (int64_t i = 0; i < plane; ++i) {
        const float t = x[o + i] - mean[c];
        out[o + i] = t / std[c];
      }
    }
}
ORC_API void orc_normalize_bwd(const float* gout, const float* std, float* gin, int B, int C, int64_t plane) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const int64_t o = ((int64_t)b * C + c) * plane;
      for (int64_t i = 0; i < plane; ++i) gin[o + i] = gout[o + i] / std[c];
    }
}

/* ---- input_transformation/sim.py:40  cat([x / 2**i]) -------------------------------------------------- */
ORC_API void orc_sim_fwd(const float* x, float* out, int S, int64_t N) {
  for (int s = 0; s < S; ++s) {
    const float d = (float)(1u << s);
    for (int64_t i = 0; i < N; ++i) out[(int64_t)s * N + i] = x[i] / d;
  }
}
/* autograd accumulates the S branch gradients into x's buffer later-created-first: s = S-1 down to 0 */
ORC_API void orc_sim_bwd(const float* gout, float* gin, int S, int64_t N) {
  for (int64_t i = 0; i < N; ++i) {
    float acc = gout[(int64_t)(S - 1) * N + i] / (float)(1u << (S - 1));
    for (int s = S - 2; s >= 0; --s) {
      const float t = gout[(int64_t)s * N + i] / (float)(1u << s);
      acc = acc + t;
    }
    gin[i] = acc;
  }
}

/* ---- input_transformation/admix.py:44-45 ------------------------------------------------------------------ */
ORC_API void orc_admix_fwd(const float* x, const int32_t* perm, float strength, float* out, int S, int A, int B,
                           int64_t n) {
  for (int s = 0; s < S; ++s)
    for (int a = 0; a < A; ++a)
      for (int b = 0; b < B; ++b) {
        const float* xs = x + (int64_t)b * n;
        const float* xp = x + (int64_t)perm[a * B + b] * n;
        float* o = out + (((int64_t)s * A + a) * B + b) * n;
        const float d = (float)(1u << s);
        for (int64_t i = 0; i < n; ++i) {
          const float t = strength * xp[i];
          const float u = xs[i] + t;
          o[i] = u / d;
        }
      }
}
ORC_API void orc_admix_bwd(const float* gout, float* gin, int S, int A, int B, int64_t n) {
  for (int b = 0; b < B; ++b)
    for (int64_t i = 0; i < n; ++i) {
      float outer = 0.0f;
      for (int a = A - 1; a >= 0; --a) {
        float acc = gout[((((int64_t)(S - 1)) * A + a) * B + b) * n + i] / (float)(1u << (S - 1));
        for (int s = S - 2; s >= 0; --s) {
          const float t = gout[(((int64_t)s * A + a) * B + b) * n + i] / (float)(1u << s);
          acc = acc + t;
        }
        outer = (a == A - 1) ? acc : outer + acc;
      }
      gin[(int64_t)b * n + i] = outer;
    }
}

/* ---- input_transformation/dim.py:55,65,68 — F.interpolate(bilinear, align_corners=False) + F.pad -------------
 * ATen upsample_bilinear2d: scale = (float)in/(float)out; src = max(0, fmaf(scale, dst + 0.5f, -0.5f));
 * i0 = (int)src; i1 = i0 + (i0 < in-1); l1 = src - i0; l0 = 1 - l1;
 * val = hl0*(wl0*p00 + wl1*p01) + hl1*(wl0*p10 + wl1*p11)                                                  */
typedef struct { int i0, i1; float l0, l1; } lin_tap;
static void lin_taps(int in, int out, lin_tap* t) {
  const float scale = (float)in / (float)out;
  for (int d = 0; d < out; ++d) {
    float src = fmaf(scale, (float)d + 0.5f, -0.5f);
    if (src < 0.0f) src = 0.0f;
    const int i0 = (int)src;
    t[d].i0 = i0;
    t[d].i1 = i0 + ((i0 < in - 1) ? 1 : 0);
    t[d].l1 = src - (float)i0;
    t[d].l0 = 1.0f - t[d].l1;
  }
}
static int g_blend_mode = 0; /* 0: every op rounded; 1-4: FMA contractions of ATen's expression (see csrc/dim.cu) */
ORC_API void orc_set_dim_blend(int mode) { g_blend_mode = mode; }
static void bilinear_plane(const float* in, int ih, int iw, float* out, int oh, int ow, const lin_tap* th,
                           const lin_tap* tw) {
  (void)ih;
  const int mode = g_blend_mode;
  for (int y = 0; y < oh; ++y) {
    const float* r0 = in + (int64_t)th[y].i0 * iw;
    const float* r1 = in + (int64_t)th[y].i1 * iw;
    for (int x = 0; x < ow; ++x) {
      const float w0 = tw[x].l0, w1 = tw[x].l1, h0 = th[y].l0, h1 = th[y].l1;
      const float p00 = r0[tw[x].i0], p01 = r0[tw[x].i1], p10 = r1[tw[x].i0], p11 = r1[tw[x].i1];
      float top, bot, val;
      if (mode == 0) {
        const float a = w0 * p00, b = w1 * p01, c = w0 * p10, d = w1 * p11;
        top = a + b; bot = c + d;
        const float e = h0 * top, f = h1 * bot;
        val = e + f;
      } else {
        if (mode == 1 || mode == 3) { const float b = w1 * p01, d = w1 * p11; top = fmaf(w0, p00, b); bot = fmaf(w0, p10, d); }
        else { const float a = w0 * p00, c = w0 * p10; top = fmaf(w1, p01, a); bot = fmaf(w1, p11, c); }
        if (mode == 1 || mode == 4) { const float f = h1 * bot; val = fmaf(h0, top, f); }
        else { const float e = h0 * top; val = fmaf(h1, bot, e); }
      }
      out[(int64_t)y * ow + x] = val;
    }
  }
}
ORC_API int orc_dim_fwd(const float* x, float* out, int planes, int S, int rnd, int R, int pad_top, int pad_left) {
  lin_tap* t1 = (lin_tap*)malloc(sizeof(lin_tap) * (size_t)rnd);
  lin_tap* t2 = (lin_tap*)malloc(sizeof(lin_tap) * (size_t)S);
  float* y1 = (float*)malloc(sizeof(float) * (size_t)rnd * rnd);
  float* y2 = (float*)malloc(sizeof(float) * (size_t)R * R);
  if (!t1 || !t2 || !y1 || !y2) return -1;
  lin_taps(S, rnd, t1); /* S -> rnd */
  lin_taps(R, S, t2);   /* R -> S */
  for (int p = 0; p < planes; ++p) {
    bilinear_plane(x + (int64_t)p * S * S, S, S, y1, rnd, rnd, t1, t1);
    memset(y2, 0, sizeof(float) * (size_t)R * R);
    for (int y = 0; y < rnd; ++y)
      memcpy(y2 + (int64_t)(y + pad_top) * R + pad_left, y1 + (int64_t)y * rnd, sizeof(float) * (size_t)rnd);
    bilinear_plane(y2, R, R, out + (int64_t)p * S * S, S, S, t2, t2);
  }
  free(t1); free(t2); free(y1); free(y2);
  return 0;
}
/* exact adjoint, fp64 scatter accumulation (order-free reference for the deterministic gather kernel;
 * ATen's own backward scatters with atomicAdd in fp32, i.e. is itself order-dependent) */
static void bilinear_plane_adj(const float* gout, int oh, int ow, double* gin, int ih, int iw, const lin_tap* th,
                               const lin_tap* tw) {
  (void)ih;
  for (int y = 0; y < oh; ++y)
    for (int x = 0; x < ow; ++x) {
      const double g = (double)gout[(int64_t)y * ow + x];
      gin[(int64_t)th[y].i0 * iw + tw[x].i0] += (double)th[y].l0 * (double)tw[x].l0 * g;
      gin[(int64_t)th[y].i0 * iw + tw[x].i1] += (double)th[y].l0 * (double)tw[x].l1 * g;
      gin[(int64_t)th[y].i1 * iw + tw[x].i0] += (double)th[y].l1 * (double)tw[x].l0 * g;
      gin[(int64_t)th[y].i1 * iw + tw[x].i1] += (double)th[y].l1 * (double)tw[x].l1 * g;
    }
}
ORC_API int orc_dim_bwd(const float* gout, float* gin, int planes, int S, int rnd, int R, int pad_top, int pad_left) {
  lin_tap* t1 = (lin_tap*)malloc(sizeof(lin_tap) * (size_t)rnd);
  lin_tap* t2 = (lin_tap*)malloc(sizeof(lin_tap) * (size_t)S);
  double* g2 = (double*)malloc(sizeof(double) * (size_t)R * R);
  float* g1 = (float*)malloc(sizeof(float) * (size_t)rnd * rnd);
  double* g0 = (double*)malloc(sizeof(double) * (size_t)S * S);
  if (!t1 || !t2 || !g2 || !g1 || !g0) return -1;
  lin_taps(S, rnd, t1);
  lin_taps(R, S, t2);
  for (int p = 0; p < planes; ++p) {
    memset(g2, 0, sizeof(double) * (size_t)R * R);
    bilinear_plane_adj(gout + (int64_t)p * S * S, S, S, g2, R, R, t2, t2);
    /* pad adjoint = crop; the intermediate gradient is an fp32 tensor in the reference */
    for (int y = 0; y < rnd; ++y)
      for (int x = 0; x < rnd; ++x) g1[(int64_t)y * rnd + x] = (float)g2[(int64_t)(y + pad_top) * R + (x + pad_left)];
    memset(g0, 0, sizeof(double) * (size_t)S * S);
    bilinear_plane_adj(g1, rnd, rnd, g0, S, S, t1, t1);
    for (int64_t i = 0; i < (int64_t)S * S; ++i) gin[(int64_t)p * S * S + i] = (float)g0[i];
  }
  free(t1); free(t2); free(g2); free(g1); free(g0);
  return 0;
}

/* ---- input_transformation/tim.py:73  F.conv2d(grad, K, padding='same', groups=C) ------------------------------
 * fp32 FMA chain in (ky, kx) raster order from 0 — the order the 2-D CUDA kernel replays bit-exactly. */
ORC_API void orc_dwconv2d(const float* g, const float* k, int ks, float* out, int B, int C, int H, int W) {
  const int r = ks / 2;
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float* gp = g + ((int64_t)b * C + c) * H * W;
      float* op = out + ((int64_t)b * C + c) * H * W;
      const float* kc = k + (int64_t)c * ks * ks;
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          float acc = 0.0f;
          for (int i = 0; i < ks; ++i) {
            const int yy = y + i - r;
            for (int j = 0; j < ks; ++j) {
              const int xx = x + j - r;
              const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? gp[(int64_t)yy * W + xx] : 0.0f;
              acc = fmaf(kc[i * ks + j], v, acc);
            }
          }
          op[(int64_t)y * W + x] = acc;
        }
    }
}
/* separable form: rows first (krow over x), then columns (kcol over y), fp32 FMA chains from 0 */
ORC_API int orc_dwconv2d_sep(const float* g, const float* kcol, const float* krow, int ks, float* out, int B, int C,
                             int H, int W) {
  const int r = ks / 2;
  float* tmp = (float*)malloc(sizeof(float) * (size_t)H * W);
  if (!tmp) return -1;
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float* gp = g + ((int64_t)b * C + c) * H * W;
      float* op = out + ((int64_t)b * C + c) * H * W;
      const float* kr = krow + (int64_t)c * ks;
      const float* kc = kcol + (int64_t)c * ks;
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          float acc = 0.0f;
          for (int j = 0; j < ks; ++j) {
            const int xx = x + j - r;
            const float v = (xx >= 0 && xx < W) ? gp[(int64_t)y * W + xx] : 0.0f;
            acc = fmaf(kr[j], v, acc);
          }
          tmp[(int64_t)y * W + x] = acc;
        }
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          float acc = 0.0f;
          for (int i = 0; i < ks; ++i) {
            const int yy = y + i - r;
            const float v = (yy >= 0 && yy < H) ? tmp[(int64_t)yy * W + x] : 0.0f;
            acc = fmaf(kc[i], v, acc);
          }
          op[(int64_t)y * W + x] = acc;
        }
    }
  free(tmp);
  return 0;
}

/* ---- gradient/emifgsm.py:57-58  concat([x + factor*alpha*grad]) ----------------------------------------------------- */
ORC_API void orc_lin_sample_fwd(const float* x, const float* gbar, const float* coef, int K, float* out, int64_t N) {
  for (int k = 0; k < K; ++k)
    for (int64_t i = 0; i < N; ++i) {
      const float t = gbar ? coef[k] * gbar[i] : 0.0f;
      out[(int64_t)k * N + i] = x[i] + t;
    }
}
ORC_API void orc_lin_sample_bwd(const float* gout, float* gin, int K, int64_t N) {
  for (int64_t i = 0; i < N; ++i) {
    float acc = gout[(int64_t)(K - 1) * N + i];
    for (int k = K - 2; k >= 0; --k) acc = acc + gout[(int64_t)k * N + i];
    gin[i] = acc;
  }
}

/* ---- gradient/vmifgsm.py:50,56,58 ---------------------------------------------------------------------------------- */
ORC_API void orc_neighbor_stage(const float* data, const float* delta, const float* noise, const float* look, float coef,
                                float* out, int64_t N) {
  for (int64_t j = 0; j < N; ++j) {
    const float x = data[j] + delta[j];
    const float xn = x + noise[j];
    if (look) {
      const float t = coef * look[j];
      out[j] = xn + t;
    } else {
      out[j] = xn;
    }
  }
}
ORC_API void orc_accumulate(float* acc, const float* g, int first, int64_t N) {
  for (int64_t j = 0; j < N; ++j) acc[j] = first ? g[j] : acc[j] + g[j];
}
ORC_API void orc_variance_finalize(const float* acc, const float* cur, int num_neighbor, float* out, int64_t N) {
  const float d = (float)num_neighbor;
  for (int64_t j = 0; j < N; ++j) {
    const float t = acc[j] / d;
    out[j] = t - cur[j];
  }
}
ORC_API void orc_add(const float* a, const float* b, float* out, int64_t N) {
  for (int64_t j = 0; j < N; ++j) out[j] = a[j] + b[j];
}

/* ---- utils.py:64  (adversaries.permute(0,2,3,1).numpy() * 255).astype(np.uint8) ----------------------------------- */
ORC_API void orc_quantize_u8(const float* data, const float* delta, uint8_t* out, int B, int C, int64_t plane, int to_nhwc) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int64_t i = 0; i < plane; ++i) {
        const int64_t j = ((int64_t)b * C + c) * plane + i;
        const float v = data[j] + delta[j];
        const float s = v * 255.0f;
        /* numpy float32 -> uint8: C cast (truncation toward zero) of an in-range value; mirror x86 behaviour
           for out-of-range by going through int32 and wrapping */
        const int32_t q = (int32_t)s;
        const int64_t o = to_nhwc ? ((int64_t)b * plane + i) * C + c : j;
        out[o] = (uint8_t)q;
      }
}

/* ---- gradient/pifgsm.py:94-102, 61-68 (PI-FGSM; project_noise itself is orc_dwconv2d) -------------------------------- */
ORC_API void orc_pi_cut_noise(const float* amp, const float* m, float coef, float eps, float* amp_out, float* cut_out, int64_t N) {
  for (int64_t j = 0; j < N; ++j) {
    const float st = coef * sgnf(m[j]);
    const float a1 = (amp ? amp[j] : 0.0f) + st;
    const float t = fabsf(a1) - eps;
    const float c = min_nan(max_nan(t, 0.0f), 10000.0f);
    amp_out[j] = a1;
    cut_out[j] = c * sgnf(a1);
  }
}
ORC_API void orc_pi_update_linf(const float* delta, const float* data, const float* g, const float* conv, const float* amp,
                                float alpha, float gamma, float eps, float lo, float hi, float* amp_out, float* delta_out,
                                int64_t N) {
  const float neg_eps = -eps;
  for (int64_t j = 0; j < N; ++j) {
    const float proj = gamma * sgnf(conv[j]);
    const float a2 = amp[j] + proj;
    const float st = alpha * sgnf(g[j]);
    const float d0 = delta[j] + st;
    const float d1 = d0 + proj;
    const float d2 = min_nan(max_nan(d1, neg_eps), eps);
    const float l = lo - data[j], h = hi - data[j];
    amp_out[j] = a2;
    delta_out[j] = min_nan(max_nan(d2, l), h);
  }
}

/* ---- gradient/gra.py:74-93 + :149 (GRA decay indicator + tensor-step update_delta) ------------------------------------- */
ORC_API void orc_gra_update(const float* M, const float* last, const float* cur, float eta, float alpha, const float* delta,
                            const float* data, float eps, float lo, float hi, float* M_out, float* delta_out, int64_t N) {
  const float neg_eps = -eps;
  for (int64_t j = 0; j < N; ++j) {
    const float sl = last ? sgnf(last[j]) : 0.0f, sc = sgnf(cur[j]);
    const float eq = (sl == sc) ? 1.0f : 0.0f;
    const float di = 1.0f - eq;
    const float t = di * eta;
    const float f = eq + t;
    const float m = M[j] * f;
    const float a = m * alpha;
    const float st = a * sc;
    const float d1 = delta[j] + st;
    const float d2 = min_nan(max_nan(d1, neg_eps), eps);
    const float l = lo - data[j], h = hi - data[j];
    M_out[j] = m;
    delta_out[j] = min_nan(max_nan(d2, l), h);
  }
}

/* ---- ensemble/adaea.py:115-136 (disparity-reduced filter) + :74-76 (threshold) + :82 (grad * mask) ------------------------
 * grads: K pointers to [B, C, plane]; map_out [B, plane] (nullable); out = grad * mask (both nullable together).           */
ORC_API void orc_adaea_drf(const float* const* grads, int K, float threshold, const float* grad, float* out, float* map_out,
                           int B, int C, int64_t plane) {
  float u[8][4];
  for (int b = 0; b < B; ++b)
    for (int64_t q = 0; q < plane; ++q) {
      for (int k = 0; k < K; ++k) {
        float s = 0.0f, w[4], s2 = 0.0f;
        for (int c = 0; c < C; ++c) { const float v = grads[k][((int64_t)b * C + c) * plane + q]; const float vv = v * v; s = s + vv; }
        const float nrm = sqrtf(s);
        const float den = nrm > 1e-12f ? nrm : 1e-12f;
        for (int c = 0; c < C; ++c) { w[c] = grads[k][((int64_t)b * C + c) * plane + q] / den; const float ww = w[c] * w[c]; s2 = s2 + ww; }
        const float n2 = sqrtf(s2);
        const float den2 = n2 > 1e-8f ? n2 : 1e-8f;
        for (int c = 0; c < C; ++c) u[k][c] = w[c] / den2;
      }
      float tot = 0.0f;
      for (int a = 0; a < K - 1; ++a) {
        float row = 0.0f;
        for (int j = 0; j < K; ++j) {
          if (j == a) continue;
          float d = 0.0f;
          for (int c = 0; c < C; ++c) { const float pr = u[a][c] * u[j][c]; d = d + pr; }
          row = row + d;
        }
        tot = tot + row / (float)(K - 1);
      }
      const float m = tot / (float)K;
      const float mask = (m >= threshold) ? 1.0f : ((m < threshold) ? 0.0f : m);
      if (map_out) map_out[(int64_t)b * plane + q] = m;
      if (grad)
        for (int c = 0; c < C; ++c) { const int64_t i = ((int64_t)b * C + c) * plane + q; out[i] = grad[i] * mask; }
    }
}
