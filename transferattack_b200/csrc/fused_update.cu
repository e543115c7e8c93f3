// fused_update.cu — ta_fused_update_linf: the whole tail of one attack iteration in ONE launch
//   (attack.py:124-128 get_momentum, :145-153 update_delta, and the next iteration's :88 `data + delta`).
//
//   mu_b   = mean|g_b|                               per sample b
//   m'     = m * decay + g / mu_b
//   delta' = clamp(clamp(delta + alpha*sign(m'), -eps, eps), lo - x, hi - x)
//   xadv   = x + delta'
//
// HBM roofline: 16 B/elem read (g, m, delta, x) + 12 B/elem written (m', delta', xadv) = 28 B/elem
// (24 without xadv). The per-sample mean needs all of g_b before the first output can be formed, so the
// kernel runs one thread-block CLUSTER per sample:
//   phase A: every CTA pulls its slice of g_b into shared memory with bulk-TMA (cp.async.bulk, chunked on
//            mbarriers so the |g| reduction of chunk c overlaps the transfer of chunk c+1), reduces it in
//            fp64, and the cluster combines the partials through DSMEM in rank order;
//   phase B: streams m, delta, x with 128-bit loads, takes g from shared memory (so g crosses HBM once),
//            and writes m', delta', xadv with 128-bit stores.
// Variant 1 keeps nothing in shared memory and re-reads g in phase B (an L2 hit: the cluster touched it
// microseconds earlier); it trades L2 bandwidth for occupancy. Both are exposed through ta_tune_set for the
// sweep in bench/; the default is chosen from measurements.
//
// With `scale` given (strict mode: torch computed mean|g| with the reference's own op) there is no phase A
// and the work is a flat 128-bit streaming kernel.
#include "common.cuh"

using namespace ta;

namespace {

// Normalize folding (reference utils.py:72-79): the model input the kernel emits is the NORMALISED image
// (x + delta' - mean_c) / std_c (torchvision's sub_ then div_: two roundings), and — when `bwd` — the incoming gradient is
// the one w.r.t. that normalised input, turned into the gradient w.r.t. delta by Normalize's adjoint g / std_c first.
struct NormFold {
  float mean[4], std[4];
  int64_t plane_vec;      // vectors (of the launch's width) per channel plane
  int C, fwd, bwd;
};

struct FusedParams {
  const float* g; const float* m; float* m_out; const float* delta; float* delta_out; const float* data;
  float* xadv; const float* scale; float* scale_out;
  float decay, alpha, eps, lo, hi;
  int64_t n;
  NormFold nf;
};

// channel of vector j (index inside one sample); C <= 4
__device__ __forceinline__ int nf_channel(const NormFold& nf, int64_t j) {
  int c = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k) c += (k < nf.C && j >= k * nf.plane_vec) ? 1 : 0;
  return c;
}

// one element of the fused tail; all roundings as in the reference's eager ops
__device__ __forceinline__ void fused_elem(float g, float m, bool has_m, float d, float x, float mu, const FusedParams& p,
                                           float& m_new, float& d_new, float& xa) {
  const float t1 = has_m ? mul_rn(m, p.decay) : 0.0f;
  m_new = add_rn(t1, div_rn(g, mu));
  d_new = project_linf(d, mul_rn(p.alpha, sign_t(m_new)), x, p.eps, p.lo, p.hi);
  xa = add_rn(x, d_new);
}
template <bool NF>
__device__ __forceinline__ void fused_elem_nf(float g, float m, bool has_m, float d, float x, float mu, const FusedParams& p,
                                              float mean_c, float std_c, float& m_new, float& d_new, float& xa) {
  if (NF && p.nf.bwd) g = div_rn(g, std_c);
  fused_elem(g, m, has_m, d, x, mu, p, m_new, d_new, xa);
  if (NF && p.nf.fwd) xa = div_rn(sub_rn(xa, mean_c), std_c);
}

// ---- strict / fallback path: scale[b] given, flat streaming ---------------------------------------------------
template <int V> struct FusedIn { Vec<V> g, x, d, m; float mu; };
template <bool NF>
struct FusedStreamOpT {
  FusedParams p; int64_t nvec;     // vectors per sample
  template <int V> __device__ __forceinline__ FusedIn<V> load(int row, int64_t j) const {
    FusedIn<V> r;
    const int64_t i = (int64_t)row * nvec + j;
    r.mu = __ldg(p.scale + row);
    r.g = ldv<V>(p.g, i); r.x = ldv<V>(p.data, i); r.d = ldv_rw<V>(p.delta, i);
    if (p.m) r.m = ldv_rw<V>(p.m, i);
    return r;
  }
  template <int V> __device__ __forceinline__ void apply(int row, int64_t j, const FusedIn<V>& r) const {
    const int64_t i = (int64_t)row * nvec + j;
    Vec<V> mo, dn, xa;
    float mean_c = 0.0f, std_c = 1.0f;
    if (NF) { const int c = nf_channel(p.nf, j); mean_c = p.nf.mean[c]; std_c = p.nf.std[c]; }
#pragma unroll
    for (int k = 0; k < V; ++k)
      fused_elem_nf<NF>(r.g.v[k], p.m ? r.m.v[k] : 0.0f, p.m != nullptr, r.d.v[k], r.x.v[k], r.mu, p, mean_c, std_c, mo.v[k],
                        dn.v[k], xa.v[k]);
    stv<V>(p.m_out, i, mo);
    stv<V>(p.delta_out, i, dn);
    if (p.xadv) stv<V>(p.xadv, i, xa);
  }
};
using FusedStreamOp = FusedStreamOpT<false>;

// ---- cluster kernel ----------------------------------------------------------------------------------------------
constexpr int kChunks = 4;

// STAGE = true : g slice resident in shared memory (bulk-TMA), read from HBM once
// STAGE = false: g re-read through L2 in phase B
__device__ __forceinline__ double abs4(const float4& v) {
  double a = (double)fabsf(v.x); a += (double)fabsf(v.y); a += (double)fabsf(v.z); a += (double)fabsf(v.w); return a;
}
template <int THREADS, int U, bool STAGE, bool NF>
__global__ void __launch_bounds__(THREADS) fused_cluster_kernel(FusedParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_scratch[32];
  __shared__ double s_part;
  __shared__ __align__(8) uint64_t s_bar[kChunks];

  const int tid = threadIdx.x;
  const int64_t nvec = p.n >> 2;
  const int64_t nr = cluster_nctarank(), rank = cluster_ctarank();
  const int64_t per = (nvec + nr - 1) / nr;
  const int64_t begin = rank * per < nvec ? rank * per : nvec;
  const int64_t end = (rank + 1) * per < nvec ? (rank + 1) * per : nvec;
  const int64_t cnt = end - begin;                                   // 128-bit vectors in this CTA's slice
  const int64_t off = (int64_t)blockIdx.y * nvec + begin;            // slice start, in vectors, in the batch
  const float4* g4 = reinterpret_cast<const float4*>(p.g) + off;
  float4* sg4 = reinterpret_cast<float4*>(smem_raw);
  const int64_t per_chunk = (cnt + kChunks - 1) / kChunks;

  // ---------------- phase A: sum |g| over the slice ----------------
  double acc = 0.0;
  if (STAGE) {
    if (tid == 0) {
#pragma unroll
      for (int c = 0; c < kChunks; ++c) mbar_init(&s_bar[c], 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int64_t c0 = c * per_chunk, c1 = (c0 + per_chunk < cnt) ? c0 + per_chunk : cnt;
        if (c1 > c0) {
          const uint32_t bytes = (uint32_t)((c1 - c0) * 16);
          mbar_expect_tx(&s_bar[c], bytes);
          tma_bulk_g2s(sg4 + c0, g4 + c0, bytes, &s_bar[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int64_t c0 = c * per_chunk, c1 = (c0 + per_chunk < cnt) ? c0 + per_chunk : cnt;
      if (c1 > c0) {
        mbar_wait(&s_bar[c], 0);
        for (int64_t i = c0 + tid; i < c1; i += THREADS) {
          float4 v = sg4[i];
          if (NF && p.nf.bwd) { const float sc = p.nf.std[nf_channel(p.nf, begin + i)]; v.x = div_rn(v.x, sc); v.y = div_rn(v.y, sc); v.z = div_rn(v.z, sc); v.w = div_rn(v.w, sc); }
          acc += (double)fabsf(v.x); acc += (double)fabsf(v.y); acc += (double)fabsf(v.z); acc += (double)fabsf(v.w);
        }
      }
    }
  } else {
    for (int64_t i = tid; i < cnt; i += THREADS) {
      float4 v = __ldg(g4 + i);
      if (NF && p.nf.bwd) { const float sc = p.nf.std[nf_channel(p.nf, begin + i)]; v.x = div_rn(v.x, sc); v.y = div_rn(v.y, sc); v.z = div_rn(v.z, sc); v.w = div_rn(v.w, sc); }
      acc += (double)fabsf(v.x); acc += (double)fabsf(v.y); acc += (double)fabsf(v.z); acc += (double)fabsf(v.w);
    }
  }
  const double part = block_sum(acc, s_scratch);
  if (tid == 0) s_part = part;
  cluster_sync_all();
  double tot = 0.0;
  for (uint32_t r = 0; r < (uint32_t)nr; ++r) tot += dsmem_ld_f64(&s_part, r);
  cluster_arrive();                       // "done reading remote shared memory"; matched by cluster_wait() at exit
  const float mu = (float)(tot / (double)p.n);
  if (rank == 0 && tid == 0 && p.scale_out) p.scale_out[blockIdx.y] = mu;

  // ---------------- phase B: stream the update ----------------
  const bool has_m = p.m != nullptr;
  const float4* m4 = reinterpret_cast<const float4*>(p.m) + off;
  const float4* d4 = reinterpret_cast<const float4*>(p.delta) + off;
  const float4* x4 = reinterpret_cast<const float4*>(p.data) + off;
  float4* mo4 = reinterpret_cast<float4*>(p.m_out) + off;
  float4* do4 = reinterpret_cast<float4*>(p.delta_out) + off;
  float4* xa4 = reinterpret_cast<float4*>(p.xadv) + off;
  for (int64_t i0 = tid; i0 < cnt; i0 += (int64_t)THREADS * U) {
    float4 gv[U], mv[U], dv[U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + (int64_t)u * THREADS;
      if (i < cnt) {
        gv[u] = STAGE ? sg4[i] : __ldg(g4 + i);
        xv[u] = __ldg(x4 + i);
        dv[u] = d4[i];
        mv[u] = has_m ? m4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + (int64_t)u * THREADS;
      if (i < cnt) {
        float4 mo, dn, xa;
        float mean_c = 0.0f, std_c = 1.0f;
        if (NF) { const int c = nf_channel(p.nf, begin + i); mean_c = p.nf.mean[c]; std_c = p.nf.std[c]; }
        fused_elem_nf<NF>(gv[u].x, mv[u].x, has_m, dv[u].x, xv[u].x, mu, p, mean_c, std_c, mo.x, dn.x, xa.x);
        fused_elem_nf<NF>(gv[u].y, mv[u].y, has_m, dv[u].y, xv[u].y, mu, p, mean_c, std_c, mo.y, dn.y, xa.y);
        fused_elem_nf<NF>(gv[u].z, mv[u].z, has_m, dv[u].z, xv[u].z, mu, p, mean_c, std_c, mo.z, dn.z, xa.z);
        fused_elem_nf<NF>(gv[u].w, mv[u].w, has_m, dv[u].w, xv[u].w, mu, p, mean_c, std_c, mo.w, dn.w, xa.w);
        mo4[i] = mo;
        do4[i] = dn;
        if (p.xadv) xa4[i] = xa;
      }
    }
  }
  cluster_wait();                         // keep s_part alive until every rank has read it
}

template <int THREADS, int U, bool STAGE, bool NF = false>
int launch_fused(const FusedParams& p, int B, int cl, size_t smem, cudaStream_t s) {
  auto k = fused_cluster_kernel<THREADS, U, STAGE, NF>;
  static SmemOptIn optin = {};
  static bool nonportable[64] = {};
  int rc = ensure_dyn_smem("ta_fused_update_linf", k, smem, optin);
  if (rc != TA_OK) return rc;
  if (cl > 8) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!nonportable[dev]) {
      const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) {
        set_error("ta_fused_update_linf: cluster size %d not allowed: %s", cl, cudaGetErrorString(e));
        cudaGetLastError();
        return TA_ECUDA;
      }
      nonportable[dev] = true;
    }
  }
  return launch_cluster("ta_fused_update_linf", k, cl, B, THREADS, smem, s, p);
}


// ---- ENS, one surrogate per GPU: reduce-scatter + fused update + all-gather in ONE kernel over NVLink peer memory ------------
// Rank r owns the samples [b0, b0 + Bown). For those samples the kernel
//   reads the K per-rank gradient buffers (K-1 of them are PEER memory mapped over NVLink) and sums them in the order
//   autograd accumulates the members' gradients on one device (k = K-1 first, then K-2, ... 0),
//   runs exactly ta_fused_update_linf's arithmetic (cluster per sample, summed g kept in shared memory),
//   and stores x_adv = x + delta' into EVERY rank's model-input buffer (K-1 remote stores per element),
// i.e. the gradient reduce-scatter, the update and the all-gather of the next model input are one launch; m' and delta' stay
// local to the owner. Cross-GPU ordering (all gradients written before / all x_adv visible after) is the caller's two
// symmetric-memory barriers on the same stream. L1 is invalidated at every kernel launch, and peer lines bypass the local L2,
// so plain loads see the peers' fresh data.
constexpr int kMaxPeers = 8;
struct PeerPtrs { const float* g[kMaxPeers]; float* x[kMaxPeers]; int K; };

template <int THREADS, int U>
__global__ void __launch_bounds__(THREADS) fused_p2p_kernel(FusedParams p, PeerPtrs pp, int b0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_scratch[32];
  __shared__ double s_part;
  const int tid = threadIdx.x;
  const int K = pp.K;
  const int64_t nvec = p.n >> 2;
  const int64_t nr = cluster_nctarank(), rank = cluster_ctarank();
  const int64_t per = (nvec + nr - 1) / nr;
  const int64_t begin = rank * per < nvec ? rank * per : nvec;
  const int64_t end = (rank + 1) * per < nvec ? (rank + 1) * per : nvec;
  const int64_t cnt = end - begin;
  const int64_t off = (int64_t)(b0 + blockIdx.y) * nvec + begin;       // slice start (vectors) in the FULL batch
  float4* sg4 = reinterpret_cast<float4*>(smem_raw);

  // phase A: g = sum over ranks (descending rank order), kept in shared memory; sum |g| in fp64
  double acc = 0.0;
  for (int64_t i0 = tid; i0 < cnt; i0 += (int64_t)THREADS * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + (int64_t)u * THREADS;
      if (i < cnt) v[u] = reinterpret_cast<const float4*>(pp.g[K - 1])[off + i];
    }
    for (int k = K - 2; k >= 0; --k) {
      float4 w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * THREADS;
        if (i < cnt) w[u] = reinterpret_cast<const float4*>(pp.g[k])[off + i];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        v[u].x = add_rn(v[u].x, w[u].x); v[u].y = add_rn(v[u].y, w[u].y); v[u].z = add_rn(v[u].z, w[u].z); v[u].w = add_rn(v[u].w, w[u].w);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + (int64_t)u * THREADS;
      if (i < cnt) {
        sg4[i] = v[u];
        acc += (double)fabsf(v[u].x); acc += (double)fabsf(v[u].y); acc += (double)fabsf(v[u].z); acc += (double)fabsf(v[u].w);
      }
    }
  }
  const double part = block_sum(acc, s_scratch);
  if (tid == 0) s_part = part;
  cluster_sync_all();
  double tot = 0.0;
  for (uint32_t r = 0; r < (uint32_t)nr; ++r) tot += dsmem_ld_f64(&s_part, r);
  cluster_arrive();
  float mu = (float)(tot / (double)p.n);
  if (p.scale) mu = __ldg(p.scale + b0 + blockIdx.y);
  if (rank == 0 && tid == 0 && p.scale_out) p.scale_out[b0 + blockIdx.y] = mu;

  // phase B
  const bool has_m = p.m != nullptr;
  const float4* m4 = reinterpret_cast<const float4*>(p.m) + off;
  const float4* d4 = reinterpret_cast<const float4*>(p.delta) + off;
  const float4* x4 = reinterpret_cast<const float4*>(p.data) + off;
  float4* mo4 = reinterpret_cast<float4*>(p.m_out) + off;
  float4* do4 = reinterpret_cast<float4*>(p.delta_out) + off;
  for (int64_t i0 = tid; i0 < cnt; i0 += (int64_t)THREADS * U) {
    float4 mv[U], dv[U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + (int64_t)u * THREADS;
      if (i < cnt) {
        xv[u] = __ldg(x4 + i);
        dv[u] = d4[i];
        mv[u] = has_m ? m4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + (int64_t)u * THREADS;
      if (i < cnt) {
        const float4 gv = sg4[i];
        float4 mo, dn, xa;
        fused_elem(gv.x, mv[u].x, has_m, dv[u].x, xv[u].x, mu, p, mo.x, dn.x, xa.x);
        fused_elem(gv.y, mv[u].y, has_m, dv[u].y, xv[u].y, mu, p, mo.y, dn.y, xa.y);
        fused_elem(gv.z, mv[u].z, has_m, dv[u].z, xv[u].z, mu, p, mo.z, dn.z, xa.z);
        fused_elem(gv.w, mv[u].w, has_m, dv[u].w, xv[u].w, mu, p, mo.w, dn.w, xa.w);
        mo4[i] = mo;
        do4[i] = dn;
        for (int k = 0; k < K; ++k) reinterpret_cast<float4*>(pp.x[k])[off + i] = xa;      // local + K-1 peers over NVLink
      }
    }
  }
  cluster_wait();
}

constexpr size_t kMaxStageBytes = 200 * 1024;   // per-CTA g slice bound (227 KB/SM minus static + system use)

}  // namespace

namespace {

int fused_update_impl(const float* g, const float* m, float* m_out, const float* delta, float* delta_out, const float* data,
                      float* xadv_out, const float* scale, float* scale_out, int mean_mode, float decay, float alpha, float eps,
                      float lo, float hi, int B, int64_t n, const NormFold* nf, ta_stream_t stream) {
  const char* who = nf ? "ta_fused_update_linf_nf" : "ta_fused_update_linf";
  TA_REQUIRE(g && m_out && delta && delta_out && data && B > 0 && n > 0, "%s: null pointer or empty shape (B=%d n=%lld)", who, B,
             (long long)n);
  TA_REQUIRE(B <= 65535, "%s: B=%d exceeds 65535", who, B);
  cudaStream_t s = (cudaStream_t)stream;
  FusedParams p{g, m, m_out, delta, delta_out, data, xadv_out, scale, scale_out, decay, alpha, eps, lo, hi, n};
  const bool v4 = (n % 4 == 0) && aligned16(g) && aligned16(m) && aligned16(m_out) && aligned16(delta) &&
                  aligned16(delta_out) && aligned16(data) && aligned16(xadv_out);
  if (nf) {
    if (!v4) { set_error("%s: needs n %% 4 == 0 and 16-byte aligned buffers", who); return TA_EUNSUPPORTED; }
    p.nf = *nf;          // plane_vec already in 128-bit vectors
  }

  if (scale) {   // strict: no reduction, flat streaming
    if (scale_out && scale_out != scale) {
      const cudaError_t e = cudaMemcpyAsync(scale_out, scale, sizeof(float) * (size_t)B, cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) { set_error("%s: scale copy failed: %s", who, cudaGetErrorString(e)); return TA_ECUDA; }
    }
    if (nf) return launch_ew_rows2<1>("ta_fused_update_linf_nf[stream]", B, n, true, FusedStreamOpT<true>{p, n / 4}, s, 0);
    const FusedStreamOp op{p, v4 ? n / 4 : n};
    const int cap = tune_get("stream.cap", 0);          // resident CTAs per SM (0 = one batch per thread, no loop)
    switch (tune_get("stream.unroll", 1)) {
      case 4: return launch_ew_rows2<4>("ta_fused_update_linf[stream]", B, n, v4, op, s, cap);
      case 2: return launch_ew_rows2<2>("ta_fused_update_linf[stream]", B, n, v4, op, s, cap);
      default: return launch_ew_rows2<1>("ta_fused_update_linf[stream]", B, n, v4, op, s, cap);
    }
  }

  if (mean_mode != TA_MEAN_EXACT) {
    set_error("%s: mean_mode %d not available in this build", who, mean_mode);
    return TA_EUNSUPPORTED;
  }

  // cluster geometry
  int cl = tune_get("fused.cluster", 0);
  if (cl <= 0) {
    cl = 1;
    while (cl < 8 && n / (cl * 2) >= 2048) cl *= 2;          // >= 2K elements per CTA before splitting further
  }
  const int variant = tune_get("fused.variant", 0);          // 0 = g staged in smem by bulk-TMA, 1 = re-read via L2
  const int threads = tune_get("fused.threads", 512);
  const int unroll = tune_get("fused.unroll", 2);
  const int64_t nvec = n / 4;
  const size_t slice_bytes = (size_t)((nvec + cl - 1) / cl) * 16;

  if (!v4) {
    // generic fallback: exact mean into scale_out, then the streaming kernel (two launches)
    TA_REQUIRE(scale_out, "ta_fused_update_linf: n %% 4 != 0 or misaligned pointers need scale_out as scratch");
    const int rc = ta_abs_mean_per_sample(g, scale_out, B, n, TA_MEAN_EXACT, nullptr, stream);
    if (rc != TA_OK) return rc;
    p.scale = scale_out;
    return launch_ew_rows2<2>("ta_fused_update_linf[stream]", B, n, false, FusedStreamOp{p, n}, s);
  }

  const bool stage = (variant == 0) && slice_bytes <= kMaxStageBytes;
  const size_t smem = stage ? slice_bytes : 0;
  if (nf)   // one tuning point (the default) for the folded form
    return stage ? launch_fused<512, 2, true, true>(p, B, cl, smem, s) : launch_fused<512, 2, false, true>(p, B, cl, 0, s);
#define TA_FUSED_CASE(T, U_)                                                        \
  if (threads == T && unroll == U_)                                                 \
    return stage ? launch_fused<T, U_, true>(p, B, cl, smem, s) : launch_fused<T, U_, false>(p, B, cl, 0, s);
  TA_FUSED_CASE(256, 1)
  TA_FUSED_CASE(256, 2)
  TA_FUSED_CASE(256, 4)
  TA_FUSED_CASE(512, 1)
  TA_FUSED_CASE(512, 2)
  TA_FUSED_CASE(512, 4)
  TA_FUSED_CASE(1024, 1)
  TA_FUSED_CASE(1024, 2)
#undef TA_FUSED_CASE
  set_error("ta_fused_update_linf: unsupported tuning threads=%d unroll=%d", threads, unroll);
  return TA_EUNSUPPORTED;
}

}  // namespace

extern "C" int ta_fused_update_linf(const float* g, const float* m, float* m_out, const float* delta, float* delta_out,
                                    const float* data, float* xadv_out, const float* scale, float* scale_out, int mean_mode,
                                    float decay, float alpha, float eps, float lo, float hi, int B, int64_t n,
                                    ta_stream_t stream) {
  return fused_update_impl(g, m, m_out, delta, delta_out, data, xadv_out, scale, scale_out, mean_mode, decay, alpha, eps, lo, hi,
                           B, n, nullptr, stream);
}

// Normalize folded in: `xn_out` receives the NORMALISED next model input ((data + delta') - mean_c) / std_c,
// channel c = (element index inside the sample) / plane; with grad_wrt_xn != 0, `g` is the gradient w.r.t. that normalised
// input and is divided by std_c (Normalize's adjoint) before anything else. mean_host / std_host: HOST arrays [C], C <= 4.
extern "C" int ta_fused_update_linf_nf(const float* g, const float* m, float* m_out, const float* delta, float* delta_out,
                                       const float* data, float* xn_out, const float* scale, float* scale_out, int mean_mode,
                                       float decay, float alpha, float eps, float lo, float hi, int B, int64_t n,
                                       const float* mean_host, const float* std_host, int C, int64_t plane, int grad_wrt_xn,
                                       ta_stream_t stream) {
  TA_REQUIRE(mean_host && std_host && xn_out, "ta_fused_update_linf_nf: null pointer");
  if (C < 1 || C > 4 || plane <= 0 || plane % 4 != 0 || (int64_t)C * plane != n) {
    set_error("ta_fused_update_linf_nf: needs 1 <= C <= 4, plane %% 4 == 0 and C * plane == n (C=%d plane=%lld n=%lld)", C,
              (long long)plane, (long long)n);
    return TA_EUNSUPPORTED;
  }
  NormFold nf = {};
  for (int c = 0; c < C; ++c) {
    TA_REQUIRE(std_host[c] != 0.0f, "ta_fused_update_linf_nf: std[%d] == 0", c);
    nf.mean[c] = mean_host[c]; nf.std[c] = std_host[c];
  }
  nf.C = C; nf.fwd = 1; nf.bwd = grad_wrt_xn ? 1 : 0; nf.plane_vec = plane / 4;
  return fused_update_impl(g, m, m_out, delta, delta_out, data, xn_out, scale, scale_out, mean_mode, decay, alpha, eps, lo, hi, B,
                           n, &nf, stream);
}

extern "C" int ta_fused_allreduce_update_linf(const float* const* g_peers, float* const* xadv_peers, int K, const float* m,
                                              float* m_out, const float* delta, float* delta_out, const float* data,
                                              const float* scale, float* scale_out, int mean_mode, float decay, float alpha,
                                              float eps, float lo, float hi, int b0, int Bown, int64_t n, ta_stream_t stream) {
  TA_REQUIRE(g_peers && xadv_peers && K >= 1 && K <= kMaxPeers, "ta_fused_allreduce_update_linf: K=%d (1..%d)", K, kMaxPeers);
  TA_REQUIRE(m_out && delta && delta_out && data && b0 >= 0 && n > 0, "ta_fused_allreduce_update_linf: null pointer or bad shape");
  if (Bown <= 0) return TA_OK;
  TA_REQUIRE(Bown <= 65535, "ta_fused_allreduce_update_linf: Bown=%d exceeds 65535", Bown);
  if (mean_mode != TA_MEAN_EXACT) { set_error("ta_fused_allreduce_update_linf: mean_mode %d not available", mean_mode); return TA_EUNSUPPORTED; }
  PeerPtrs pp;
  pp.K = K;
  bool ok = (n % 4 == 0) && aligned16(m) && aligned16(m_out) && aligned16(delta) && aligned16(delta_out) && aligned16(data);
  for (int k = 0; k < kMaxPeers; ++k) {
    pp.g[k] = k < K ? g_peers[k] : nullptr;
    pp.x[k] = k < K ? xadv_peers[k] : nullptr;
    if (k < K) { TA_REQUIRE(pp.g[k] && pp.x[k], "ta_fused_allreduce_update_linf: null peer pointer %d", k); ok = ok && aligned16(pp.g[k]) && aligned16(pp.x[k]); }
  }
  TA_REQUIRE(ok, "ta_fused_allreduce_update_linf: needs n %% 4 == 0 and 16-byte aligned buffers");
  FusedParams p{nullptr, m, m_out, delta, delta_out, data, nullptr, scale, scale_out, decay, alpha, eps, lo, hi, n};
  const int64_t nvec = n / 4;
  int cl = tune_get("fused.cluster", 0);
  if (cl <= 0) { cl = 1; while (cl < 8 && n / (cl * 2) >= 2048) cl *= 2; }
  size_t slice = (size_t)((nvec + cl - 1) / cl) * 16;
  while (slice > kMaxStageBytes && cl < 16) { cl *= 2; slice = (size_t)((nvec + cl - 1) / cl) * 16; }
  if (slice > kMaxStageBytes) { set_error("ta_fused_allreduce_update_linf: sample of %lld elements does not fit a 16-CTA cluster", (long long)n); return TA_EUNSUPPORTED; }
  auto k = fused_p2p_kernel<512, 2>;
  static SmemOptIn optin = {};
  static bool nonportable[64] = {};
  int rc = ensure_dyn_smem("ta_fused_allreduce_update_linf", k, slice, optin);
  if (rc != TA_OK) return rc;
  if (cl > 8) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!nonportable[dev]) {
      if (cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { set_error("cluster 16 not allowed"); cudaGetLastError(); return TA_ECUDA; }
      nonportable[dev] = true;
    }
  }
  return launch_cluster("ta_fused_allreduce_update_linf", k, cl, Bown, 512, slice, (cudaStream_t)stream, p, pp, b0);
}
