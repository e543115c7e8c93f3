// elementwise.cu — the HBM-bound streaming kernels behind the Attack hooks: momentum, L-inf update,
// box clamp, model-input staging, Normalize, SIM / Admix / EMI replication + adjoints, VMI neighbour ops,
// uint8 quantisation. All are 128-bit vectorised grid-stride kernels (scalar fallback when a pointer or the
// per-sample length is not 16-byte friendly); one rounding per reference op (see common.cuh).
#include "common.cuh"

using namespace ta;

namespace {

template <class... P>
bool all_aligned(P... p) {
  bool ok = true;
  const void* a[] = {static_cast<const void*>(p)...};
  for (const void* q : a) ok = ok && (q == nullptr || aligned16(q));
  return ok;
}

// ---- attack.py:128 ---------------------------------------------------------------------------------------
template <int V> struct MomIn { Vec<V> g, m; float mu; };
struct MomentumOp {
  const float* g; const float* m; const float* scale; float* out; float decay; int64_t nvec;   // nvec: vectors per sample
  template <int V> __device__ __forceinline__ MomIn<V> load(int row, int64_t j) const {
    MomIn<V> r;
    const int64_t i = (int64_t)row * nvec + j;
    r.mu = __ldg(scale + row);
    r.g = ldv<V>(g, i);
    if (m) r.m = ldv_rw<V>(m, i);
    return r;
  }
  template <int V> __device__ __forceinline__ void apply(int row, int64_t j, const MomIn<V>& r) const {
    Vec<V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) o.v[k] = add_rn(m ? mul_rn(r.m.v[k], decay) : 0.0f, div_rn(r.g.v[k], r.mu));
    stv<V>(out, (int64_t)row * nvec + j, o);
  }
};

// ---- attack.py:147,152 -------------------------------------------------------------------------------------
template <int V> struct UpdIn { Vec<V> d, x, g, a; };
struct UpdateLinfOp {
  const float* delta; const float* data; const float* dir; const float* alpha_t; float* out;
  float alpha, eps, lo, hi; int dir_mode;
  template <int V> __device__ __forceinline__ UpdIn<V> load(int64_t i) const {
    UpdIn<V> r;
    r.d = ldv_rw<V>(delta, i); r.x = ldv<V>(data, i); r.g = ldv<V>(dir, i);
    if (alpha_t) r.a = ldv<V>(alpha_t, i);
    return r;
  }
  template <int V> __device__ __forceinline__ void apply(int64_t i, const UpdIn<V>& r) const {
    Vec<V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float d = (dir_mode == TA_DIR_SIGN) ? sign_t(r.g.v[k]) : r.g.v[k];
      const float a = alpha_t ? r.a.v[k] : alpha;
      o.v[k] = project_linf(r.d.v[k], mul_rn(a, d), r.x.v[k], eps, lo, hi);
    }
    stv<V>(out, i, o);
  }
};

// ---- attack.py:141 -------------------------------------------------------------------------------------------
struct ClampBoxOp {
  const float* delta; const float* data; float* out; float lo, hi;
  template <int V> __device__ void run(int64_t i) const {
    const Vec<V> dv = ldv_rw<V>(delta, i), xv = ldv<V>(data, i);
    Vec<V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) o.v[k] = min_nan(max_nan(dv.v[k], sub_rn(lo, xv.v[k])), sub_rn(hi, xv.v[k]));
    stv<V>(out, i, o);
  }
};

// ---- gradient/pifgsm.py:94-98 (PI-FGSM) ------------------------------------------------------------------------
//   amplification += (beta*alpha) * sign(momentum)
//   cut_noise      = clamp(|amplification| - eps, 0, 10000) * sign(amplification)
struct PiCutOp {
  const float* amp; const float* m; float* amp_out; float* cut_out; float coef, eps;
  template <int V> __device__ void run(int64_t i) const {
    const Vec<V> mv = ldv<V>(m, i);
    Vec<V> av, ao, co;
    if (amp) av = ldv_rw<V>(amp, i);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float a1 = add_rn(amp ? av.v[k] : 0.0f, mul_rn(coef, sign_t(mv.v[k])));
      const float c = min_nan(max_nan(sub_rn(fabsf(a1), eps), 0.0f), 10000.0f);
      ao.v[k] = a1;
      co.v[k] = mul_rn(c, sign_t(a1));
    }
    stv<V>(amp_out, i, ao);
    stv<V>(cut_out, i, co);
  }
};
//   projection     = gamma * sign(conv3x3(cut_noise));  amplification += projection                       (pifgsm.py:99-100)
//   delta          = clamp(delta + alpha*sign(g) + projection, -eps, eps), then the box clamp              (pifgsm.py:61-68)
struct PiUpdateOp {
  const float* delta; const float* data; const float* g; const float* conv; const float* amp; float* amp_out; float* delta_out;
  float alpha, gamma, eps, lo, hi;
  template <int V> __device__ void run(int64_t i) const {
    const Vec<V> dv = ldv_rw<V>(delta, i), xv = ldv<V>(data, i), gv = ldv<V>(g, i), cv = ldv<V>(conv, i), av = ldv_rw<V>(amp, i);
    Vec<V> ao, dn;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float proj = mul_rn(gamma, sign_t(cv.v[k]));
      ao.v[k] = add_rn(av.v[k], proj);
      const float d1 = add_rn(add_rn(dv.v[k], mul_rn(alpha, sign_t(gv.v[k]))), proj);
      const float d2 = min_nan(max_nan(d1, -eps), eps);
      dn.v[k] = min_nan(max_nan(d2, sub_rn(lo, xv.v[k])), sub_rn(hi, xv.v[k]));
    }
    stv<V>(amp_out, i, ao);
    stv<V>(delta_out, i, dn);
  }
};

// ---- attack.py:88 / nifgsm.py:39 / vmifgsm.py:50 ------------------------------------------------------------------
template <int V> struct StageIn { Vec<V> x, d, n, l; };
struct StageOp {
  const float* data; const float* delta; const float* noise; const float* look; float* out; float coef;
  template <int V> __device__ __forceinline__ StageIn<V> load(int64_t i) const {
    StageIn<V> r;
    r.x = ldv<V>(data, i);
    if (delta) r.d = ldv<V>(delta, i);
    if (noise) r.n = ldv<V>(noise, i);
    if (look) r.l = ldv<V>(look, i);
    return r;
  }
  template <int V> __device__ __forceinline__ void apply(int64_t i, const StageIn<V>& r) const {
    Vec<V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float x = delta ? add_rn(r.x.v[k], r.d.v[k]) : r.x.v[k];
      if (noise) x = add_rn(x, r.n.v[k]);
      if (look) x = add_rn(x, mul_rn(coef, r.l.v[k]));
      o.v[k] = x;
    }
    stv<V>(out, i, o);
  }
};

// ---- utils.py:72-79 Normalize ---------------------------------------------------------------------------------------
struct NormalizeOp {
  const float* x; const float* mean; const float* std; float* out; int C; int64_t nvec; bool fwd;    // nvec: vectors per plane
  template <int V> __device__ void run(int row, int64_t j) const {     // row = b * C + c
    const int c = row % C;
    const int64_t i = (int64_t)row * nvec + j;
    const float sd = __ldg(std + c);
    const Vec<V> xv = ldv_rw<V>(x, i);
    Vec<V> o;
    if (fwd) {
      const float mu = __ldg(mean + c);
#pragma unroll
      for (int k = 0; k < V; ++k) o.v[k] = div_rn(sub_rn(xv.v[k], mu), sd);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) o.v[k] = div_rn(xv.v[k], sd);
    }
    stv<V>(out, i, o);
  }
};

// ---- sim.py:40 ----------------------------------------------------------------------------------------------------------
struct SimFwdOp {
  const float* x; float* out; int S; int64_t nvec_per_copy;
  template <int V> __device__ void run(int64_t i) const {
    const Vec<V> xv = ldv<V>(x, i);
    for (int s = 0; s < S; ++s) {
      const float d = (float)(1u << s);
      Vec<V> o;
#pragma unroll
      for (int k = 0; k < V; ++k) o.v[k] = div_rn(xv.v[k], d);
      stv<V>(out, (int64_t)s * nvec_per_copy + i, o);
    }
  }
};
struct SimBwdOp {
  const float* gout; float* gin; int S; int64_t nvec_per_copy;
  template <int V> __device__ void run(int64_t i) const {
    Vec<V> acc = ldv<V>(gout, (int64_t)(S - 1) * nvec_per_copy + i);
    {
      const float d = (float)(1u << (S - 1));
#pragma unroll
      for (int k = 0; k < V; ++k) acc.v[k] = div_rn(acc.v[k], d);
    }
    for (int s = S - 2; s >= 0; --s) {
      const Vec<V> gv = ldv<V>(gout, (int64_t)s * nvec_per_copy + i);
      const float d = (float)(1u << s);
#pragma unroll
      for (int k = 0; k < V; ++k) acc.v[k] = add_rn(acc.v[k], div_rn(gv.v[k], d));
    }
    stv<V>(gin, i, acc);
  }
};

// ---- admix.py:44-45 ------------------------------------------------------------------------------------------------------
struct AdmixFwdOp {
  const float* x; const int32_t* perm; float* out; float strength; int S, A, B; int64_t nv;   // nv = vectors per sample
  template <int V> __device__ void run(int64_t i) const {   // i over A*B*nv
    const int64_t e = i % nv;
    const int64_t ab = i / nv;
    const int b = (int)(ab % B);
    const int src = __ldg(perm + ab);
    const Vec<V> xs = ldv<V>(x, (int64_t)b * nv + e), xp = ldv<V>(x, (int64_t)src * nv + e);
    Vec<V> u;
#pragma unroll
    for (int k = 0; k < V; ++k) u.v[k] = add_rn(xs.v[k], mul_rn(strength, xp.v[k]));
    for (int s = 0; s < S; ++s) {
      const float d = (float)(1u << s);
      Vec<V> o;
#pragma unroll
      for (int k = 0; k < V; ++k) o.v[k] = div_rn(u.v[k], d);
      stv<V>(out, ((int64_t)s * A * B + ab) * nv + e, o);
    }
  }
};
struct AdmixBwdOp {
  const float* gout; float* gin; int S, A, B; int64_t nv;
  template <int V> __device__ void run(int64_t i) const {   // i over B*nv
    const int64_t e = i % nv;
    const int b = (int)(i / nv);
    Vec<V> outer;
    for (int a = A - 1; a >= 0; --a) {
      Vec<V> acc = ldv<V>(gout, (((int64_t)(S - 1) * A + a) * B + b) * nv + e);
      {
        const float d = (float)(1u << (S - 1));
#pragma unroll
        for (int k = 0; k < V; ++k) acc.v[k] = div_rn(acc.v[k], d);
      }
      for (int s = S - 2; s >= 0; --s) {
        const Vec<V> gv = ldv<V>(gout, (((int64_t)s * A + a) * B + b) * nv + e);
        const float d = (float)(1u << s);
#pragma unroll
        for (int k = 0; k < V; ++k) acc.v[k] = add_rn(acc.v[k], div_rn(gv.v[k], d));
      }
      if (a == A - 1) outer = acc;
      else {
#pragma unroll
        for (int k = 0; k < V; ++k) outer.v[k] = add_rn(outer.v[k], acc.v[k]);
      }
    }
    stv<V>(gin, i, outer);
  }
};

// ---- emifgsm.py:57-58 ----------------------------------------------------------------------------------------------------------
struct CoefTable { float c[32]; };
struct LinSampleFwdOp {
  const float* x; const float* gbar; float* out; CoefTable coef; int K; int64_t nvec_per_copy;
  template <int V> __device__ void run(int64_t i) const {
    const Vec<V> xv = ldv<V>(x, i);
    Vec<V> gv;
    if (gbar) gv = ldv<V>(gbar, i);
    for (int k = 0; k < K; ++k) {
      Vec<V> o;
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = add_rn(xv.v[j], gbar ? mul_rn(coef.c[k], gv.v[j]) : 0.0f);
      stv<V>(out, (int64_t)k * nvec_per_copy + i, o);
    }
  }
};
struct LinSampleBwdOp {
  const float* gout; float* gin; int K; int64_t nvec_per_copy;
  template <int V> __device__ void run(int64_t i) const {
    Vec<V> acc = ldv<V>(gout, (int64_t)(K - 1) * nvec_per_copy + i);
    for (int k = K - 2; k >= 0; --k) {
      const Vec<V> gv = ldv<V>(gout, (int64_t)k * nvec_per_copy + i);
#pragma unroll
      for (int j = 0; j < V; ++j) acc.v[j] = add_rn(acc.v[j], gv.v[j]);
    }
    stv<V>(gin, i, acc);
  }
};

// ---- vmifgsm.py:56,58,87 ----------------------------------------------------------------------------------------------------------
template <int V> struct AccIn { Vec<V> g, a; };
struct AccumulateOp {
  float* acc; const float* g; int first;
  template <int V> __device__ __forceinline__ AccIn<V> load(int64_t i) const {
    AccIn<V> r;
    r.g = ldv<V>(g, i);
    if (!first) r.a = ldv_rw<V>(acc, i);
    return r;
  }
  template <int V> __device__ __forceinline__ void apply(int64_t i, const AccIn<V>& r) const {
    Vec<V> o = r.g;
    if (!first) {
#pragma unroll
      for (int k = 0; k < V; ++k) o.v[k] = add_rn(r.a.v[k], r.g.v[k]);
    }
    stv<V>(acc, i, o);
  }
};
struct VarianceOp {
  const float* acc; const float* cur; float* out; float nn;
  template <int V> __device__ void run(int64_t i) const {
    const Vec<V> av = ldv_rw<V>(acc, i), cv = ldv_rw<V>(cur, i);
    Vec<V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) o.v[k] = sub_rn(div_rn(av.v[k], nn), cv.v[k]);
    stv<V>(out, i, o);
  }
};
struct AddOp {
  const float* a; const float* b; float* out;
  template <int V> __device__ void run(int64_t i) const {
    const Vec<V> av = ldv_rw<V>(a, i), bv = ldv_rw<V>(b, i);
    Vec<V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) o.v[k] = add_rn(av.v[k], bv.v[k]);
    stv<V>(out, i, o);
  }
};

// ---- utils.py:64 save_images quantisation ---------------------------------------------------------------------------------------------
// One thread per (b, pixel): reads C planes coalesced across the warp, writes C consecutive bytes.
__global__ void __launch_bounds__(256) quantize_kernel(const float* __restrict__ data, const float* __restrict__ delta,
                                                       uint8_t* __restrict__ out, int B, int C, int64_t plane, int to_nhwc) {
  const int64_t total = (int64_t)B * plane;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t b = t / plane, i = t % plane;
    for (int c = 0; c < C; ++c) {
      const int64_t j = (b * C + c) * plane + i;
      const float v = mul_rn(add_rn(__ldg(data + j), __ldg(delta + j)), 255.0f);
      const int q = __float2int_rz(v);                       // numpy astype(uint8): truncation toward zero
      out[to_nhwc ? (b * plane + i) * C + c : j] = (uint8_t)q;
    }
  }
}

// vectorised: each thread takes 4 consecutive pixels of one image (float4 per channel) and, for C == 3 NHWC, writes the
// 12 output bytes as three 32-bit stores. grid = (x, B).
__global__ void __launch_bounds__(256) quantize_nhwc3_kernel(const float* __restrict__ data, const float* __restrict__ delta,
                                                             uint8_t* __restrict__ out, int64_t plane4) {
  const int b = blockIdx.y;
  const float4* x4 = reinterpret_cast<const float4*>(data) + (int64_t)b * 3 * plane4;
  const float4* d4 = reinterpret_cast<const float4*>(delta) + (int64_t)b * 3 * plane4;
  uint32_t* o = reinterpret_cast<uint32_t*>(out) + (int64_t)b * 3 * plane4;      // 12 bytes = 3 words per 4 pixels
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane4; i += stride) {
    uint32_t q[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 xv = __ldg(x4 + c * plane4 + i), dv = __ldg(d4 + c * plane4 + i);
      q[c][0] = (uint32_t)(uint8_t)__float2int_rz(mul_rn(add_rn(xv.x, dv.x), 255.0f));
      q[c][1] = (uint32_t)(uint8_t)__float2int_rz(mul_rn(add_rn(xv.y, dv.y), 255.0f));
      q[c][2] = (uint32_t)(uint8_t)__float2int_rz(mul_rn(add_rn(xv.z, dv.z), 255.0f));
      q[c][3] = (uint32_t)(uint8_t)__float2int_rz(mul_rn(add_rn(xv.w, dv.w), 255.0f));
    }
    // bytes in memory order: p0c0 p0c1 p0c2 p1c0 | p1c1 p1c2 p2c0 p2c1 | p2c2 p3c0 p3c1 p3c2
    o[3 * i + 0] = q[0][0] | (q[1][0] << 8) | (q[2][0] << 16) | (q[0][1] << 24);
    o[3 * i + 1] = q[1][1] | (q[2][1] << 8) | (q[0][2] << 16) | (q[1][2] << 24);
    o[3 * i + 2] = q[2][2] | (q[0][3] << 8) | (q[1][3] << 16) | (q[2][3] << 24);
  }
}

}  // namespace

extern "C" {

int ta_momentum(const float* g, const float* m, const float* scale, float decay, float* m_out, int B, int64_t n,
                ta_stream_t stream) {
  TA_REQUIRE(g && scale && m_out && B > 0 && n > 0, "ta_momentum: null pointer or empty shape (B=%d n=%lld)", B, (long long)n);
  const bool v4 = (n % 4 == 0) && all_aligned(g, m, m_out);
  return launch_ew_rows2<1>("ta_momentum", B, n, v4, MomentumOp{g, m, scale, m_out, decay, v4 ? n / 4 : n}, (cudaStream_t)stream);
}

int ta_update_linf(const float* delta, const float* data, const float* dir, const float* alpha_t, float alpha, float eps,
                   float lo, float hi, int dir_mode, float* delta_out, int64_t N, ta_stream_t stream) {
  TA_REQUIRE(delta && data && dir && delta_out && N > 0, "ta_update_linf: null pointer or N=%lld", (long long)N);
  TA_REQUIRE(dir_mode == TA_DIR_SIGN || dir_mode == TA_DIR_RAW, "ta_update_linf: dir_mode %d", dir_mode);
  const bool v4 = (N % 4 == 0) && all_aligned(delta, data, dir, alpha_t, delta_out);
  return launch_ew2<1>("ta_update_linf", N, v4, UpdateLinfOp{delta, data, dir, alpha_t, delta_out, alpha, eps, lo, hi, dir_mode},
                   (cudaStream_t)stream);
}

int ta_clamp_box(const float* delta, const float* data, float lo, float hi, float* out, int64_t N, ta_stream_t stream) {
  TA_REQUIRE(delta && data && out && N > 0, "ta_clamp_box: null pointer or N=%lld", (long long)N);
  const bool v4 = (N % 4 == 0) && all_aligned(delta, data, out);
  return launch_ew("ta_clamp_box", N, v4, ClampBoxOp{delta, data, out, lo, hi}, (cudaStream_t)stream);
}

int ta_stage_add(const float* data, const float* delta, const float* look, float coef, float* out, int64_t N,
                 ta_stream_t stream) {
  TA_REQUIRE(data && out && N > 0, "ta_stage_add: null pointer or N=%lld", (long long)N);
  const bool v4 = (N % 4 == 0) && all_aligned(data, delta, look, out);
  return launch_ew2<1>("ta_stage_add", N, v4, StageOp{data, delta, nullptr, look, out, coef}, (cudaStream_t)stream);
}

int ta_pi_cut_noise(const float* amp, const float* momentum, float coef, float eps, float* amp_out, float* cut_out, int64_t N,
                    ta_stream_t stream) {
  TA_REQUIRE(momentum && amp_out && cut_out && N > 0, "ta_pi_cut_noise: null pointer or N=%lld", (long long)N);
  const bool v4 = (N % 4 == 0) && all_aligned(amp, momentum, amp_out, cut_out);
  return launch_ew("ta_pi_cut_noise", N, v4, PiCutOp{amp, momentum, amp_out, cut_out, coef, eps}, (cudaStream_t)stream);
}

int ta_pi_update_linf(const float* delta, const float* data, const float* g, const float* conv, const float* amp, float alpha,
                      float gamma, float eps, float lo, float hi, float* amp_out, float* delta_out, int64_t N, ta_stream_t stream) {
  TA_REQUIRE(delta && data && g && conv && amp && amp_out && delta_out && N > 0, "ta_pi_update_linf: null pointer or N=%lld",
             (long long)N);
  const bool v4 = (N % 4 == 0) && all_aligned(delta, data, g, conv, amp, amp_out, delta_out);
  return launch_ew("ta_pi_update_linf", N, v4, PiUpdateOp{delta, data, g, conv, amp, amp_out, delta_out, alpha, gamma, eps, lo, hi},
                   (cudaStream_t)stream);
}

int ta_neighbor_stage(const float* data, const float* delta, const float* noise, const float* look, float coef, float* out,
                      int64_t N, ta_stream_t stream) {
  TA_REQUIRE(data && delta && noise && out && N > 0, "ta_neighbor_stage: null pointer or N=%lld", (long long)N);
  const bool v4 = (N % 4 == 0) && all_aligned(data, delta, noise, look, out);
  return launch_ew2<1>("ta_neighbor_stage", N, v4, StageOp{data, delta, noise, look, out, coef}, (cudaStream_t)stream);
}

int ta_normalize_fwd(const float* x, const float* mean, const float* std, float* out, int B, int C, int64_t plane,
                     ta_stream_t stream) {
  TA_REQUIRE(x && mean && std && out && B > 0 && C > 0 && plane > 0, "ta_normalize_fwd: bad arguments");
  const bool v4 = (plane % 4 == 0) && all_aligned(x, out);
  return launch_ew_rows("ta_normalize_fwd", B * C, plane, v4, NormalizeOp{x, mean, std, out, C, v4 ? plane / 4 : plane, true},
                        (cudaStream_t)stream);
}

int ta_normalize_bwd(const float* gout, const float* std, float* gin, int B, int C, int64_t plane, ta_stream_t stream) {
  TA_REQUIRE(gout && std && gin && B > 0 && C > 0 && plane > 0, "ta_normalize_bwd: bad arguments");
  const bool v4 = (plane % 4 == 0) && all_aligned(gout, gin);
  return launch_ew_rows("ta_normalize_bwd", B * C, plane, v4, NormalizeOp{gout, nullptr, std, gin, C, v4 ? plane / 4 : plane, false},
                        (cudaStream_t)stream);
}

int ta_sim_fwd(const float* x, float* out, int S, int64_t N, ta_stream_t stream) {
  TA_REQUIRE(x && out && S >= 1 && S <= 31 && N > 0, "ta_sim_fwd: bad arguments (S=%d N=%lld)", S, (long long)N);
  const bool v4 = (N % 4 == 0) && all_aligned(x, out);
  return launch_ew("ta_sim_fwd", N, v4, SimFwdOp{x, out, S, v4 ? N / 4 : N}, (cudaStream_t)stream);
}

int ta_sim_bwd(const float* gout, float* gin, int S, int64_t N, ta_stream_t stream) {
  TA_REQUIRE(gout && gin && S >= 1 && S <= 31 && N > 0, "ta_sim_bwd: bad arguments (S=%d N=%lld)", S, (long long)N);
  const bool v4 = (N % 4 == 0) && all_aligned(gout, gin);
  return launch_ew("ta_sim_bwd", N, v4, SimBwdOp{gout, gin, S, v4 ? N / 4 : N}, (cudaStream_t)stream);
}

int ta_admix_fwd(const float* x, const int32_t* perm, float strength, float* out, int S, int A, int B, int64_t n,
                 ta_stream_t stream) {
  TA_REQUIRE(x && perm && out && S >= 1 && S <= 31 && A >= 1 && B >= 1 && n > 0, "ta_admix_fwd: bad arguments");
  const bool v4 = (n % 4 == 0) && all_aligned(x, out);
  const int64_t nv = v4 ? n / 4 : n;
  return launch_ew("ta_admix_fwd", (int64_t)A * B * n, v4, AdmixFwdOp{x, perm, out, strength, S, A, B, nv}, (cudaStream_t)stream);
}

int ta_admix_bwd(const float* gout, float* gin, int S, int A, int B, int64_t n, ta_stream_t stream) {
  TA_REQUIRE(gout && gin && S >= 1 && S <= 31 && A >= 1 && B >= 1 && n > 0, "ta_admix_bwd: bad arguments");
  const bool v4 = (n % 4 == 0) && all_aligned(gout, gin);
  const int64_t nv = v4 ? n / 4 : n;
  return launch_ew("ta_admix_bwd", (int64_t)B * n, v4, AdmixBwdOp{gout, gin, S, A, B, nv}, (cudaStream_t)stream);
}

int ta_lin_sample_fwd(const float* x, const float* gbar, const float* coef_host, int K, float* out, int64_t N,
                      ta_stream_t stream) {
  TA_REQUIRE(x && coef_host && out && K >= 1 && K <= 32 && N > 0, "ta_lin_sample_fwd: bad arguments (K=%d)", K);
  CoefTable t;
  for (int k = 0; k < 32; ++k) t.c[k] = k < K ? coef_host[k] : 0.0f;
  const bool v4 = (N % 4 == 0) && all_aligned(x, gbar, out);
  return launch_ew("ta_lin_sample_fwd", N, v4, LinSampleFwdOp{x, gbar, out, t, K, v4 ? N / 4 : N}, (cudaStream_t)stream);
}

int ta_lin_sample_bwd(const float* gout, float* gin, int K, int64_t N, ta_stream_t stream) {
  TA_REQUIRE(gout && gin && K >= 1 && N > 0, "ta_lin_sample_bwd: bad arguments (K=%d)", K);
  const bool v4 = (N % 4 == 0) && all_aligned(gout, gin);
  return launch_ew("ta_lin_sample_bwd", N, v4, LinSampleBwdOp{gout, gin, K, v4 ? N / 4 : N}, (cudaStream_t)stream);
}

int ta_accumulate(float* acc, const float* g, int first, int64_t N, ta_stream_t stream) {
  TA_REQUIRE(acc && g && N > 0, "ta_accumulate: bad arguments");
  const bool v4 = (N % 4 == 0) && all_aligned(acc, g);
  return launch_ew2<1>("ta_accumulate", N, v4, AccumulateOp{acc, g, first}, (cudaStream_t)stream);
}

int ta_variance_finalize(const float* acc, const float* cur, int num_neighbor, float* out, int64_t N, ta_stream_t stream) {
  TA_REQUIRE(acc && cur && out && num_neighbor > 0 && N > 0, "ta_variance_finalize: bad arguments");
  const bool v4 = (N % 4 == 0) && all_aligned(acc, cur, out);
  return launch_ew("ta_variance_finalize", N, v4, VarianceOp{acc, cur, out, (float)num_neighbor}, (cudaStream_t)stream);
}

int ta_add(const float* a, const float* b, float* out, int64_t N, ta_stream_t stream) {
  TA_REQUIRE(a && b && out && N > 0, "ta_add: bad arguments");
  const bool v4 = (N % 4 == 0) && all_aligned(a, b, out);
  return launch_ew("ta_add", N, v4, AddOp{a, b, out}, (cudaStream_t)stream);
}

int ta_quantize_u8(const float* data, const float* delta, uint8_t* out, int B, int C, int64_t plane, int to_nhwc,
                   ta_stream_t stream) {
  TA_REQUIRE(data && delta && out && B > 0 && C > 0 && plane > 0, "ta_quantize_u8: bad arguments");
  if (C == 3 && to_nhwc && plane % 4 == 0 && all_aligned(data, delta, out) && B <= 65535) {
    const int64_t plane4 = plane / 4, want = (plane4 + 255) / 256;
    int64_t per = ((int64_t)sm_count() * 8 + B - 1) / B;
    if (per < 1) per = 1;
    dim3 grid((unsigned)(want < per ? want : per), (unsigned)B);
    quantize_nhwc3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(data, delta, out, plane4);
    count_launch();
    return check_launch("ta_quantize_u8");
  }
  const int64_t total = (int64_t)B * plane;
  const int64_t want = (total + 255) / 256, cap = (int64_t)sm_count() * 8;
  quantize_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(data, delta, out, B, C, plane, to_nhwc);
  count_launch();
  return check_launch("ta_quantize_u8");
}


// *counter = (set_to >= 0) ? set_to : *counter + delta — the device-side iteration index of loops replayed from a CUDA graph
__global__ void counter_kernel(int* counter, int delta, int set_to) { *counter = set_to >= 0 ? set_to : *counter + delta; }
int ta_counter_add(int* counter, int delta, int set_to, ta_stream_t stream) {
  TA_REQUIRE(counter, "ta_counter_add: null pointer");
  counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta, set_to);
  count_launch();
  return check_launch("ta_counter_add");
}

}  // extern "C"
