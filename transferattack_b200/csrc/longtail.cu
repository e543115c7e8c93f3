// longtail.cu — per-iteration ops of plugins outside the five BASELINE configs that many attacks reuse:
//   GRA / FGSRA decay-indicator update       gradient/gra.py:74-93 + :148-149
//   AdaEA disparity-reduced filter           ensemble/adaea.py:115-136 + :72-80
// Both are HBM-bound streaming kernels: every operand is read once with 128-bit accesses, the reference's dozens of
// elementwise / small-reduction launches become one.
#include "common.cuh"

using namespace ta;

namespace {

template <class... P>
bool all_aligned16(P... p) {
  bool ok = true;
  const void* a[] = {static_cast<const void*>(p)...};
  for (const void* q : a) ok = ok && (q == nullptr || aligned16(q));
  return ok;
}

// ---- GRA ---------------------------------------------------------------------------------------------------------------
//   last = sign(last_noise), cur = sign(cur_noise)             (gra.py:87-88; last_noise == python 0 on the first iteration)
//   eq = float(last == cur); di = 1 - eq; M' = M * (eq + di * eta)                                        (gra.py:89-91)
//   delta' = update_delta(delta, data, cur_noise, M' * alpha)                               (gra.py:149 → attack.py:145-153)
// 20 B/elem read (M, last, cur, delta, data) + 8 B/elem written (M', delta'), one launch for the reference's 17.
template <int V> struct GraIn { Vec<V> M, last, cur, d, x; };
struct GraOp {
  const float* M; const float* last; const float* cur; const float* delta; const float* data; float* M_out; float* delta_out;
  float eta, alpha, eps, lo, hi;
  template <int V> __device__ __forceinline__ GraIn<V> load(int64_t i) const {
    GraIn<V> r;
    r.M = ldv_rw<V>(M, i); r.cur = ldv<V>(cur, i); r.d = ldv_rw<V>(delta, i); r.x = ldv<V>(data, i);
    if (last) r.last = ldv<V>(last, i);
    return r;
  }
  template <int V> __device__ __forceinline__ void apply(int64_t i, const GraIn<V>& r) const {
    Vec<V> mo, dn;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float sl = last ? sign_t(r.last.v[k]) : 0.0f;
      const float sc = sign_t(r.cur.v[k]);
      const float eq = (sl == sc) ? 1.0f : 0.0f;
      const float f = add_rn(eq, mul_rn(sub_rn(1.0f, eq), eta));
      const float m = mul_rn(r.M.v[k], f);
      mo.v[k] = m;
      dn.v[k] = project_linf(r.d.v[k], mul_rn(mul_rn(m, alpha), sc), r.x.v[k], eps, lo, hi);
    }
    stv<V>(M_out, i, mo);
    stv<V>(delta_out, i, dn);
  }
};

// ---- AdaEA DRF ---------------------------------------------------------------------------------------------------------------
// Per pixel p = (b, h, w), with g_k[p] the C-vector (C <= 4) of member k's input gradient:
//   u_k = g_k / max(||g_k||_2, 1e-12)                              F.normalize(grads[k], dim=1)                (adaea.py:129)
//   cos(i, j) = <u_i / max(||u_i||, 1e-8), u_j / max(||u_j||, 1e-8)>   nn.CosineSimilarity(dim=1, eps=1e-8)      (adaea.py:124)
//   r_i = (sum_{j > i} cos(i, j) + sum_{j < i} cos(j, i)) / (K - 1)  for i < K - 1,   r_{K-1} = 0
//         (adaea.py:130-132: the `if i < j` uses the inner loop's last j = K-1, so the last member gets no row)
//   map = mean_i r_i ;  mask = map >= threshold ? 1 : (map < threshold ? 0 : map)                       (adaea.py:134, 74-76)
//   out[c] = grad[c] * mask                                                                                    (adaea.py:82)
// Threads own 4 adjacent pixels (128-bit accesses per channel plane). Sum order: channels 0..C-1, pairs in (i, j) order —
// torch's own order inside its small reductions is not specified, so parity is stated as a tolerance on `map` (tests).
constexpr int kMaxMembers = 8;
struct DrfPtrs { const float* g[kMaxMembers]; };

template <int V, int C>
__global__ void __launch_bounds__(256) adaea_drf_kernel(DrfPtrs gp, int K, const float* __restrict__ grad, float* __restrict__ out,
                                                        float* __restrict__ map_out, float threshold, int64_t plane_vec, int64_t total_vec) {
  // total_vec = B * plane_vec vectors of V pixels; channel c of pixel vector (b, q) sits at ((b*C + c) * plane_vec + q)
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int64_t b = i / plane_vec, q = i - b * plane_vec;
  const int64_t base = b * C * plane_vec + q;
  float u[kMaxMembers][C][V];
#pragma unroll
  for (int k = 0; k < kMaxMembers; ++k) {
    if (k < K) {
      Vec<V> gv[C];
#pragma unroll
      for (int c = 0; c < C; ++c) gv[c] = ldv<V>(gp.g[k], base + c * plane_vec);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) s = add_rn(s, mul_rn(gv[c].v[v], gv[c].v[v]));
        const float den = fmaxf(sqrtf(s), 1e-12f);
        float w[C], s2 = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) { w[c] = div_rn(gv[c].v[v], den); s2 = add_rn(s2, mul_rn(w[c], w[c])); }
        const float den2 = fmaxf(sqrtf(s2), 1e-8f);
#pragma unroll
        for (int c = 0; c < C; ++c) u[k][c][v] = div_rn(w[c], den2);
      }
    }
  }
  float mask[V];
  Vec<V> mp;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    float tot = 0.0f;
#pragma unroll
    for (int a = 0; a < kMaxMembers; ++a) {
      if (a < K - 1) {
        float row = 0.0f;
#pragma unroll
        for (int j = 0; j < kMaxMembers; ++j) {
          if (j < K && j != a) {
            float d = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) d = add_rn(d, mul_rn(u[a][c][v], u[j][c][v]));
            row = add_rn(row, d);
          }
        }
        tot = add_rn(tot, div_rn(row, (float)(K - 1)));
      }
    }
    const float m = div_rn(tot, (float)K);
    mp.v[v] = m;
    mask[v] = (m >= threshold) ? 1.0f : ((m < threshold) ? 0.0f : m);
  }
  if (map_out) stv<V>(map_out, i, mp);
  if (grad) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      Vec<V> gv = ldv<V>(grad, base + c * plane_vec), o;
#pragma unroll
      for (int v = 0; v < V; ++v) o.v[v] = mul_rn(gv.v[v], mask[v]);
      stv<V>(out, base + c * plane_vec, o);
    }
  }
}

}  // namespace

extern "C" {

int ta_gra_update(const float* M, const float* last, const float* cur, float eta, float alpha, const float* delta, const float* data,
                  float eps, float lo, float hi, float* M_out, float* delta_out, int64_t N, ta_stream_t stream) {
  TA_REQUIRE(M && cur && delta && data && M_out && delta_out && N > 0, "ta_gra_update: null pointer or N=%lld", (long long)N);
  const bool v4 = (N % 4 == 0) && all_aligned16(M, last, cur, delta, data, M_out, delta_out);
  return launch_ew2<1>("ta_gra_update", N, v4, GraOp{M, last, cur, delta, data, M_out, delta_out, eta, alpha, eps, lo, hi},
                       (cudaStream_t)stream);
}

int ta_adaea_drf(const float* const* grads, int K, float threshold, const float* grad, float* out, float* map_out, int B, int C,
                 int64_t plane, ta_stream_t stream) {
  TA_REQUIRE(grads && K >= 2 && K <= kMaxMembers && B > 0 && plane > 0, "ta_adaea_drf: K=%d (2..%d), B=%d, plane=%lld", K, kMaxMembers, B,
             (long long)plane);
  TA_REQUIRE((grad == nullptr) == (out == nullptr) && (grad || map_out), "ta_adaea_drf: grad and out go together; nothing to write");
  if (C != 3 && C != 1) { set_error("ta_adaea_drf: C=%d (1 or 3 supported)", C); return TA_EUNSUPPORTED; }
  DrfPtrs gp;
  bool v4 = (plane % 4 == 0) && all_aligned16(grad, out, map_out);
  for (int k = 0; k < kMaxMembers; ++k) {
    gp.g[k] = k < K ? grads[k] : nullptr;
    if (k < K) { TA_REQUIRE(gp.g[k], "ta_adaea_drf: null gradient %d", k); v4 = v4 && aligned16(gp.g[k]); }
  }
  const int V = v4 ? 4 : 1;
  const int64_t plane_vec = plane / V, total = (int64_t)B * plane_vec;
  const int64_t blocks = (total + 255) / 256;
  TA_REQUIRE(blocks <= 0x7fffffff, "ta_adaea_drf: too many pixels");
  cudaStream_t s = (cudaStream_t)stream;
  if (C == 3) {
    if (v4) adaea_drf_kernel<4, 3><<<(unsigned)blocks, 256, 0, s>>>(gp, K, grad, out, map_out, threshold, plane_vec, total);
    else adaea_drf_kernel<1, 3><<<(unsigned)blocks, 256, 0, s>>>(gp, K, grad, out, map_out, threshold, plane_vec, total);
  } else {
    if (v4) adaea_drf_kernel<4, 1><<<(unsigned)blocks, 256, 0, s>>>(gp, K, grad, out, map_out, threshold, plane_vec, total);
    else adaea_drf_kernel<1, 1><<<(unsigned)blocks, 256, 0, s>>>(gp, K, grad, out, map_out, threshold, plane_vec, total);
  }
  count_launch();
  return check_launch("ta_adaea_drf");
}

}  // extern "C"
