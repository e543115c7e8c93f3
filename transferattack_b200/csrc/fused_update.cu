// fused_update.cu — the whole tail of one attack iteration in ONE launch
//   (attack.py:124-128 get_momentum, :145-153 update_delta, and the next iteration's :88 `data + delta`).
//
//   g'     = g [/ std_c] [+ addend]                  (Normalize's adjoint when folded; VMI's `grad + variance`, vmifgsm.py:87)
//   mu_b   = mean|g'_b|                              per sample b
//   m'     = m * decay + g' / mu_b
//   delta' = clamp(clamp(delta + alpha*sign(m'), -eps, eps), lo - x, hi - x)
//   xadv   = x + delta'   [then (xadv - mean_c) / std_c when Normalize is folded]
//   gbar   = g' / mu_b                               (optional output: EMI's bar_grad, emifgsm.py:97)
//
// HBM roofline: 16 B/elem read (g, m, delta, x) + 12 B/elem written (m', delta', xadv) = 28 B/elem
// (24 without xadv). The per-sample mean needs all of g_b before the first output can be formed, so the
// kernel runs one thread-block CLUSTER per sample:
//   phase A: every CTA pulls its part of g_b into shared memory with bulk-TMA (cp.async.bulk, row groups on mbarriers so the
//            |g| reduction of group q overlaps the transfer of group q+1), reduces it, and the cluster combines through DSMEM:
//              TA_MEAN_EXACT — fp64 partial sums, combined in rank order;
//              TA_MEAN_TORCH — the fp32 summation tree of torch's own CUDA mean kernel (aten_mean.cuh), bit for bit;
//   phase B: streams m, delta, x with 128-bit loads, takes g' from shared memory (so g crosses HBM once),
//            and writes m', delta', xadv with 128-bit stores.
// Layout: the sample is viewed as rows of S 128-bit vectors; CTA r owns vector columns [r*W4, (r+1)*W4) of every row (TORCH:
// S = ATen's block threads * ctas_per_output virtual threads, so a column is one virtual thread's vectors; EXACT: S chosen
// for 16 rows).
//
// With `scale` given (torch computed mean|g| with the reference's own op) there is no phase A and the work is a flat 128-bit
// streaming kernel.
#include "aten_mean.cuh"

using namespace ta;

namespace ta { int aten_abs_mean_launch(const float* g, float* mean_out, int B, int64_t n, const MeanPre* pre, cudaStream_t s); }

namespace {

// Normalize folding (reference utils.py:72-79): the model input the kernel emits is the NORMALISED image
// (x + delta' - mean_c) / std_c (torchvision's sub_ then div_: two roundings), and — when `bwd` — the incoming gradient is
// the one w.r.t. that normalised input, turned into the gradient w.r.t. delta by Normalize's adjoint g / std_c first.
struct NormFold {
  float mean[4], std[4];
  int64_t plane_vec;      // 128-bit vectors per channel plane
  int C, fwd, bwd;
};

struct FusedParams {
  const float* g; const float* addend; const float* m; float* m_out; const float* delta; float* delta_out; const float* data;
  float* xadv; float* gbar; const float* scale; float* scale_out;
  float decay, alpha, eps, lo, hi;
  int64_t n;
  NormFold nf;
};

// channel of vector j (128-bit vector index inside one sample, j < C * plane_vec): three compares, no table
__device__ __forceinline__ int nf_channel(const NormFold& nf, int64_t j) {
  return (j >= nf.plane_vec ? 1 : 0) + (j >= 2 * nf.plane_vec ? 1 : 0) + (j >= 3 * nf.plane_vec ? 1 : 0);
}

// one element of the fused tail; all roundings as in the reference's eager ops
__device__ __forceinline__ void fused_elem(float g, float m, bool has_m, float d, float x, float mu, const FusedParams& p,
                                           float& m_new, float& d_new, float& xa, float& gb) {
  gb = div_rn(g, mu);
  const float t1 = has_m ? mul_rn(m, p.decay) : 0.0f;
  m_new = add_rn(t1, gb);
  d_new = project_linf(d, mul_rn(p.alpha, sign_t(m_new)), x, p.eps, p.lo, p.hi);
  xa = add_rn(x, d_new);
}
__device__ __forceinline__ void fused_elem(float g, float m, bool has_m, float d, float x, float mu, const FusedParams& p,
                                           float& m_new, float& d_new, float& xa) {
  float gb;
  fused_elem(g, m, has_m, d, x, mu, p, m_new, d_new, xa, gb);
}

// ---- strict / fallback path: scale[b] given, flat streaming ---------------------------------------------------
template <int V> struct FusedIn { Vec<V> g, a, x, d, m; float mu; };
template <bool NF>
struct FusedStreamOpT {
  FusedParams p; int64_t nvec;     // vectors per sample
  template <int V> __device__ __forceinline__ FusedIn<V> load(int row, int64_t j) const {
    FusedIn<V> r;
    const int64_t i = (int64_t)row * nvec + j;
    r.mu = __ldg(p.scale + row);
    r.g = ldv<V>(p.g, i); r.x = ldv<V>(p.data, i); r.d = ldv_rw<V>(p.delta, i);
    if (p.addend) r.a = ldv<V>(p.addend, i);
    if (p.m) r.m = ldv_rw<V>(p.m, i);
    return r;
  }
  template <int V> __device__ __forceinline__ void apply(int row, int64_t j, const FusedIn<V>& r) const {
    const int64_t i = (int64_t)row * nvec + j;
    Vec<V> mo, dn, xa, gb;
    float mean_c = 0.0f, std_c = 1.0f;
    if (NF) {                               // j < 2^31 (one sample): 32-bit compares
      const int jj = (int)j, pv = (int)p.nf.plane_vec;
      const int c = (jj >= pv ? 1 : 0) + (jj >= 2 * pv ? 1 : 0) + (jj >= 3 * pv ? 1 : 0);
      mean_c = pick4(p.nf.mean, c); std_c = pick4(p.nf.std, c);
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float g = r.g.v[k];
      if (NF && p.nf.bwd) g = div_rn(g, std_c);
      if (p.addend) g = add_rn(g, r.a.v[k]);
      fused_elem(g, p.m ? r.m.v[k] : 0.0f, p.m != nullptr, r.d.v[k], r.x.v[k], r.mu, p, mo.v[k], dn.v[k], xa.v[k], gb.v[k]);
      if (NF && p.nf.fwd) xa.v[k] = div_rn(sub_rn(xa.v[k], mean_c), std_c);
    }
    stv<V>(p.m_out, i, mo);
    stv<V>(p.delta_out, i, dn);
    if (p.xadv) stv<V>(p.xadv, i, xa);
    if (p.gbar) stv<V>(p.gbar, i, gb);
  }
};
using FusedStreamOp = FusedStreamOpT<false>;

// ---- cluster kernel ----------------------------------------------------------------------------------------------
constexpr int kChunks = 4;             // mbarrier-tracked row groups of the g transfer
constexpr int kThreads = kAtenThreads; // 512: ATen's block size (the TORCH tree maps one thread per block position)

struct TailLayout {
  int S4, W4;              // 128-bit vectors per row of the sample / per row of one CTA
  int rows_per_group;      // rows per mbarrier group
  unsigned long long w4_magic;   // floor(2^32 / W4) + 1: i / W4 == (i * magic) >> 32 for the i that occur
};


// MEAN: 0 = TA_MEAN_EXACT, 1 = TA_MEAN_TORCH.  grid = (cluster, B), 512 threads.
// dynamic smem: g' part of this CTA: rows x W4 float4
template <int U, int MEAN, bool NF>
__global__ void __launch_bounds__(kThreads, (U <= 2 ? 2 : 1)) fused_cluster_kernel(FusedParams p, TailLayout L, AtenMeanCfg c) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_scratch[32];
  __shared__ double s_part;
  __shared__ float s_val[MEAN == 1 ? kAtenMaxW : 1];
  __shared__ float s_row[MEAN == 1 ? kAtenThreads : 1];
  __shared__ float s_blk[MEAN == 1 ? kAtenThreads : 1];
  __shared__ __align__(8) uint64_t s_bar[kChunks];

  const int tid = threadIdx.x;
  const int64_t nvec = p.n >> 2;
  const int rank = (int)cluster_ctarank();
  const int64_t col0 = (int64_t)rank * L.W4;                         // first vector column of this CTA
  // row j of this CTA = vectors [j*S4 + col0, +W4) of the sample, clipped to nvec: Jf full rows, then `last` vectors
  int Jf = 0;
  if (nvec >= col0 + L.W4) Jf = (int)((nvec - col0 - L.W4) / L.S4) + 1;
  const int64_t rem = nvec - ((int64_t)Jf * L.S4 + col0);
  const int last = rem > 0 ? (int)rem : 0;                            // < W4 by construction of Jf
  const int Jtot = Jf + (last > 0 ? 1 : 0);
  const int cnt = Jf * L.W4 + last;                                   // vectors of this CTA, contiguous in shared memory
  const int RG = L.rows_per_group;
  const int64_t sbase = (int64_t)blockIdx.y * nvec;                   // sample start, in vectors
  const float4* g4 = reinterpret_cast<const float4*>(p.g) + sbase;
  float4* sg4 = reinterpret_cast<float4*>(smem_raw);
  const bool fill = p.addend != nullptr;                              // g' built by the threads instead of bulk-TMA
  const bool nfb = NF && p.nf.bwd;

  // ---------------- phase A.1: g' of this CTA into shared memory ----------------
  if (!fill) {
    if (tid == 0) {
#pragma unroll
      for (int q = 0; q < kChunks; ++q) mbar_init(&s_bar[q], 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (tid < 32) {
      if (tid == 0) {
#pragma unroll
        for (int q = 0; q < kChunks; ++q) {
          const int a = q * RG, b = (a + RG < Jtot) ? a + RG : Jtot;
          if (b > a) {
            const int full = ((b < Jf ? b : Jf) - a) > 0 ? (b < Jf ? b : Jf) - a : 0;
            const int part = (last > 0 && Jf >= a && Jf < b) ? last : 0;
            mbar_expect_tx(&s_bar[q], (uint32_t)(full * L.W4 + part) * 16u);
          }
        }
      }
      __syncwarp();
      for (int j = tid; j < Jtot; j += 32) {
        const int w = j < Jf ? L.W4 : last;
        tma_bulk_g2s(sg4 + (int64_t)j * L.W4, g4 + (int64_t)j * L.S4 + col0, (uint32_t)w * 16u, &s_bar[j / RG]);
      }
    }
  } else {
    const float4* a4 = reinterpret_cast<const float4*>(p.addend) + sbase;
    for (int i = tid; i < cnt; i += kThreads) {
      const int row = (int)(((unsigned long long)i * L.w4_magic) >> 32);
      const int64_t gi = (int64_t)row * L.S4 + col0 + (i - row * L.W4);
      float4 v = __ldg(g4 + gi);
      if (nfb) v = div4(v, pick4(p.nf.std, nf_channel(p.nf, gi)));
      sg4[i] = add4(v, __ldg(a4 + gi));
    }
    __syncthreads();
  }
  const bool nfb_pass = nfb && !fill;                                 // Normalize's adjoint still to be applied (in place) below

  // ---------------- phase A.2: mean|g'| ----------------
  float mu;
  if (MEAN == 1) {
    // one thread per vector column: its virtual thread's vectors are rows 0, 1, 2, ... of that column; the 4 accumulators are
    // the 4 components (ATen's input_vectorized_thread_reduce_impl)
    for (int col = tid; col < L.W4; col += kThreads) {
      const int rows = Jf + (col < last ? 1 : 0);
      ColAcc A;
#pragma unroll 1
      for (int q = 0; q < kChunks; ++q) {
        const int a = q * RG, b = (a + RG < rows) ? a + RG : rows;
        if (b > a) {
          if (!fill) mbar_wait(&s_bar[q], 0);
          for (int j = a; j < b; ++j) {
            float4 v = sg4[j * L.W4 + col];
            if (nfb_pass) {
              v = div4(v, pick4(p.nf.std, nf_channel(p.nf, (int64_t)j * L.S4 + col0 + col)));
              sg4[j * L.W4 + col] = v;
            }
            aten_column_add(A, v);
          }
        }
      }
      s_val[col] = aten_column_value(A);
    }
    if (!fill) {                          // every thread reads rows of every group in phase B
#pragma unroll
      for (int q = 0; q < kChunks; ++q) if (q * RG < Jtot) mbar_wait(&s_bar[q], 0);
    }
    cluster_sync_all();
    mu = aten_tree_mean(c, s_val, s_row, s_blk);
    cluster_arrive();                     // "done reading remote shared memory"; matched by cluster_wait() at exit
  } else {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < kChunks; ++q) {
      const int a = q * RG, b = (a + RG < Jtot) ? a + RG : Jtot;
      if (b > a) {
        if (!fill) mbar_wait(&s_bar[q], 0);
        const int i1 = (b * L.W4 < cnt) ? b * L.W4 : cnt;
        for (int i = a * L.W4 + tid; i < i1; i += kThreads) {
          float4 v = sg4[i];
          if (nfb_pass) {
            const int row = (int)(((unsigned long long)i * L.w4_magic) >> 32);
            v = div4(v, pick4(p.nf.std, nf_channel(p.nf, (int64_t)row * L.S4 + col0 + (i - row * L.W4))));
            sg4[i] = v;
          }
          acc += (double)fabsf(v.x); acc += (double)fabsf(v.y); acc += (double)fabsf(v.z); acc += (double)fabsf(v.w);
        }
      }
    }
    const double part = block_sum(acc, s_scratch);
    if (tid == 0) s_part = part;
    cluster_sync_all();
    double tot = 0.0;
    const uint32_t nr = cluster_nctarank();
    for (uint32_t r = 0; r < nr; ++r) tot += dsmem_ld_f64(&s_part, r);
    cluster_arrive();
    mu = (float)(tot / (double)p.n);
  }
  if (rank == 0 && tid == 0 && p.scale_out) p.scale_out[blockIdx.y] = mu;
  if (nfb_pass) __syncthreads();          // in-place g / std written by other threads of this CTA (TORCH: other columns)

  // ---------------- phase B: stream the update ----------------
  const bool has_m = p.m != nullptr;
  const float4* m4 = reinterpret_cast<const float4*>(p.m) + sbase;
  const float4* d4 = reinterpret_cast<const float4*>(p.delta) + sbase;
  const float4* x4 = reinterpret_cast<const float4*>(p.data) + sbase;
  float4* mo4 = reinterpret_cast<float4*>(p.m_out) + sbase;
  float4* do4 = reinterpret_cast<float4*>(p.delta_out) + sbase;
  float4* xa4 = reinterpret_cast<float4*>(p.xadv) + sbase;
  float4* gb4 = reinterpret_cast<float4*>(p.gbar) + sbase;
  for (int i0 = tid; i0 < cnt; i0 += kThreads * U) {
    float4 mv[U], dv[U], xv[U];
    int gi[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * kThreads;
      if (i < cnt) {
        const int row = (int)(((unsigned long long)i * L.w4_magic) >> 32);
        gi[u] = row * L.S4 + (int)col0 + (i - row * L.W4);
        xv[u] = __ldg(x4 + gi[u]);
        dv[u] = d4[gi[u]];
        mv[u] = has_m ? m4[gi[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * kThreads;
      if (i < cnt) {
        const float4 gv = sg4[i];
        float4 mo, dn, xa, gb;
        fused_elem(gv.x, mv[u].x, has_m, dv[u].x, xv[u].x, mu, p, mo.x, dn.x, xa.x, gb.x);
        fused_elem(gv.y, mv[u].y, has_m, dv[u].y, xv[u].y, mu, p, mo.y, dn.y, xa.y, gb.y);
        fused_elem(gv.z, mv[u].z, has_m, dv[u].z, xv[u].z, mu, p, mo.z, dn.z, xa.z, gb.z);
        fused_elem(gv.w, mv[u].w, has_m, dv[u].w, xv[u].w, mu, p, mo.w, dn.w, xa.w, gb.w);
        if (NF && p.nf.fwd) {
          const int ch = nf_channel(p.nf, gi[u]);
          const float mean_c = pick4(p.nf.mean, ch), std_c = pick4(p.nf.std, ch);
          xa.x = div_rn(sub_rn(xa.x, mean_c), std_c); xa.y = div_rn(sub_rn(xa.y, mean_c), std_c);
          xa.z = div_rn(sub_rn(xa.z, mean_c), std_c); xa.w = div_rn(sub_rn(xa.w, mean_c), std_c);
        }
        mo4[gi[u]] = mo;
        do4[gi[u]] = dn;
        if (p.xadv) xa4[gi[u]] = xa;
        if (p.gbar) gb4[gi[u]] = gb;
      }
    }
  }
  cluster_wait();                         // keep s_part / s_val alive until every rank has read them
}

constexpr size_t kMaxStageBytes = 200 * 1024;   // per-CTA dynamic shared memory bound (227 KB/SM minus static + system use)

template <int U, int MEAN, bool NF>
int launch_fused(const char* who, const FusedParams& p, const TailLayout& L, const AtenMeanCfg& c, int B, int cl, size_t smem,
                 cudaStream_t s) {
  auto k = fused_cluster_kernel<U, MEAN, NF>;
  static SmemOptIn optin = {};
  static bool nonportable[64] = {};
  int rc = ensure_dyn_smem(who, k, smem, optin);
  if (rc != TA_OK) return rc;
  if (cl > 8) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!nonportable[dev]) {
      const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) {
        set_error("%s: cluster size %d not allowed: %s", who, cl, cudaGetErrorString(e));
        cudaGetLastError();
        return TA_ECUDA;
      }
      nonportable[dev] = true;
    }
  }
  return launch_cluster(who, k, cl, B, kThreads, smem, s, p, L, c);
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

}  // namespace

namespace {

int fused_tail_impl(const ta_fused_tail_args& a, const NormFold* nf, ta_stream_t stream) {
  const char* who = nf ? "ta_fused_tail[nf]" : "ta_fused_tail";
  const int B = a.B; const int64_t n = a.n;
  TA_REQUIRE(a.g && a.m_out && a.delta && a.delta_out && a.data && B > 0 && n > 0, "%s: null pointer or empty shape (B=%d n=%lld)", who, B,
             (long long)n);
  TA_REQUIRE(B <= 65535, "%s: B=%d exceeds 65535", who, B);
  cudaStream_t s = (cudaStream_t)stream;
  FusedParams p = {};
  p.g = a.g; p.addend = a.addend; p.m = a.m; p.m_out = a.m_out; p.delta = a.delta; p.delta_out = a.delta_out; p.data = a.data;
  p.xadv = a.xadv_out; p.gbar = a.gbar_out; p.scale = a.scale; p.scale_out = a.scale_out;
  p.decay = a.decay; p.alpha = a.alpha; p.eps = a.eps; p.lo = a.lo; p.hi = a.hi; p.n = n;
  const bool v4 = (n % 4 == 0) && aligned16(a.g) && aligned16(a.addend) && aligned16(a.m) && aligned16(a.m_out) && aligned16(a.delta) &&
                  aligned16(a.delta_out) && aligned16(a.data) && aligned16(a.xadv_out) && aligned16(a.gbar_out);
  if (nf) {
    if (!v4) { set_error("%s: needs n %% 4 == 0 and 16-byte aligned buffers", who); return TA_EUNSUPPORTED; }
    p.nf = *nf;          // plane_vec already in 128-bit vectors
  }

  if (a.scale) {   // scale given: no reduction, flat streaming
    if (a.scale_out && a.scale_out != a.scale) {
      const cudaError_t e = cudaMemcpyAsync(a.scale_out, a.scale, sizeof(float) * (size_t)B, cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) { set_error("%s: scale copy failed: %s", who, cudaGetErrorString(e)); return TA_ECUDA; }
    }
    if (nf) return launch_ew_rows2<1>("ta_fused_tail[stream,nf]", B, n, true, FusedStreamOpT<true>{p, n / 4}, s, 0);
    const FusedStreamOp op{p, v4 ? n / 4 : n};
    const int cap = tune_get("stream.cap", 0);          // resident CTAs per SM (0 = one batch per thread, no loop)
    switch (tune_get("stream.unroll", 1)) {
      case 4: return launch_ew_rows2<4>("ta_fused_tail[stream]", B, n, v4, op, s, cap);
      case 2: return launch_ew_rows2<2>("ta_fused_tail[stream]", B, n, v4, op, s, cap);
      default: return launch_ew_rows2<1>("ta_fused_tail[stream]", B, n, v4, op, s, cap);
    }
  }

  if (a.mean_mode != TA_MEAN_EXACT && a.mean_mode != TA_MEAN_TORCH) {
    set_error("%s: mean_mode %d not available in this build", who, a.mean_mode);
    return TA_EUNSUPPORTED;
  }
  const bool torch_order = a.mean_mode == TA_MEAN_TORCH;

  // Strategy for the torch-order mean. The cluster kernel reads g from HBM once but serialises "load g -> column sums ->
  // cluster barrier -> trees -> stream" per sample (two waves of 33 clusters at B = 64: measured 72 us). When the whole
  // gradient fits L2 comfortably it is faster as two launches: the mean kernel streams g once (leaving it in the 126 MB L2),
  // then the flat streaming kernel runs at full width with g as an L2 hit (measured). fused.strategy: 0 = by
  // size, 1 = always the cluster kernel, 2 = always split.
  {
    const int strategy = tune_get("fused.strategy", 0);
    const bool fits_l2 = (int64_t)B * n * 4 <= (int64_t)64 * 1024 * 1024;
    if (torch_order && v4 && a.scale_out && (strategy == 2 || (strategy == 0 && fits_l2))) {
      MeanPre pre = {};
      pre.addend = a.addend;
      if (nf && nf->bwd) { for (int c = 0; c < 4; ++c) pre.std[c] = nf->std[c]; pre.plane_vec = nf->plane_vec; }
      const int rc = aten_abs_mean_launch(a.g, a.scale_out, B, n, &pre, s);
      if (rc == TA_OK) {
        p.scale = a.scale_out;
        if (nf) return launch_ew_rows2<1>("ta_fused_tail[stream,nf]", B, n, true, FusedStreamOpT<true>{p, n / 4}, s, 0);
        return launch_ew_rows2<1>("ta_fused_tail[stream]", B, n, true, FusedStreamOp{p, n / 4}, s, 0);
      }
      if (rc != TA_EUNSUPPORTED) return rc;
    }
  }

  // cluster geometry
  int cl = tune_get("fused.cluster", 0);
  if (cl <= 0) {
    cl = 1;
    while (cl < 8 && n / (cl * 2) >= 2048) cl *= 2;          // >= 2K elements per CTA before splitting further
  }
  const int unroll = tune_get("fused.unroll", 2);
  const int64_t nvec = n / 4;
  AtenMeanCfg c = {};
  TailLayout L = {};
  size_t smem = 0;
  bool staged = v4 && nvec < (int64_t)1 << 28;
  if (staged) {
    if (torch_order) {
      const int rc = aten_mean_plan(who, B, n, cl, &c);
      if (rc != TA_OK) return rc;
      L.S4 = c.S; L.W4 = c.W4;                       // a row = ATen's S virtual threads, one 128-bit vector each
    } else {
      int64_t w4 = (nvec + (int64_t)cl * 16 - 1) / ((int64_t)cl * 16);     // ~16 rows: 4 transfer groups of 4 rows
      if (w4 < 1) w4 = 1;
      L.W4 = (int)w4; L.S4 = (int)(w4 * cl);
    }
    const int64_t rows = (nvec + L.S4 - 1) / L.S4;
    L.rows_per_group = (int)((rows + kChunks - 1) / kChunks);
    L.w4_magic = (1ull << 32) / (unsigned long long)L.W4 + 1ull;
    smem = (size_t)rows * (size_t)L.W4 * 16;
    if (smem > kMaxStageBytes || rows * L.W4 > 0x3fffffff) staged = false;
  }

  if (!staged) {
    // generic fallback (odd n, misaligned pointers, sample too large for the cluster's shared memory):
    // the mean into scale_out, then the streaming kernel (two launches)
    if (a.addend) { set_error("%s: the addend form needs n %% 4 == 0, aligned buffers and a sample that fits the cluster", who); return TA_EUNSUPPORTED; }
    if (nf && nf->bwd) { set_error("%s: grad_wrt_xn needs the staged form", who); return TA_EUNSUPPORTED; }
    TA_REQUIRE(a.scale_out, "%s: n %% 4 != 0, misaligned pointers or oversized samples need scale_out as scratch", who);
    const int rc = ta_abs_mean_per_sample(a.g, a.scale_out, B, n, a.mean_mode, nullptr, stream);
    if (rc != TA_OK) return rc;
    p.scale = a.scale_out;
    if (nf) return launch_ew_rows2<1>("ta_fused_tail[stream,nf]", B, n, true, FusedStreamOpT<true>{p, n / 4}, s, 0);
    return launch_ew_rows2<2>("ta_fused_tail[stream]", B, n, v4, FusedStreamOp{p, v4 ? n / 4 : n}, s);
  }

#define TA_FUSED_CASE(U_)                                                                                        \
  if (unroll == U_) {                                                                                            \
    if (torch_order) return nf ? launch_fused<U_, 1, true>(who, p, L, c, B, cl, smem, s)              \
                               : launch_fused<U_, 1, false>(who, p, L, c, B, cl, smem, s);            \
    return nf ? launch_fused<U_, 0, true>(who, p, L, c, B, cl, smem, s)                               \
              : launch_fused<U_, 0, false>(who, p, L, c, B, cl, smem, s);                             \
  }
  TA_FUSED_CASE(1)
  TA_FUSED_CASE(2)
  TA_FUSED_CASE(4)
#undef TA_FUSED_CASE
  set_error("%s: unsupported tuning unroll=%d", who, unroll);
  return TA_EUNSUPPORTED;
}

int make_normfold(const char* who, const float* mean_host, const float* std_host, int C, int64_t plane, int64_t n, int emit, int bwd,
                  NormFold* nf) {
  TA_REQUIRE(mean_host && std_host, "%s: null mean/std", who);
  if (C < 1 || C > 4 || plane <= 0 || plane % 4 != 0 || (int64_t)C * plane != n) {
    set_error("%s: needs 1 <= C <= 4, plane %% 4 == 0 and C * plane == n (C=%d plane=%lld n=%lld)", who, C, (long long)plane, (long long)n);
    return TA_EUNSUPPORTED;
  }
  *nf = NormFold{};
  for (int c = 0; c < C; ++c) {
    TA_REQUIRE(std_host[c] != 0.0f, "%s: std[%d] == 0", who, c);
    nf->mean[c] = mean_host[c]; nf->std[c] = std_host[c];
  }
  nf->C = C; nf->fwd = emit ? 1 : 0; nf->bwd = bwd ? 1 : 0; nf->plane_vec = plane / 4;
  return TA_OK;
}

}  // namespace

extern "C" int ta_fused_tail(const ta_fused_tail_args* a, ta_stream_t stream) {
  TA_REQUIRE(a != nullptr, "ta_fused_tail: null argument block");
  if (a->emit_normalized || a->grad_wrt_xn) {
    TA_REQUIRE(a->xadv_out || !a->emit_normalized, "ta_fused_tail: emit_normalized needs xadv_out");
    NormFold nf;
    const int rc = make_normfold("ta_fused_tail", a->mean_host, a->std_host, a->C, a->plane, a->n, a->emit_normalized, a->grad_wrt_xn, &nf);
    if (rc != TA_OK) return rc;
    return fused_tail_impl(*a, &nf, stream);
  }
  return fused_tail_impl(*a, nullptr, stream);
}

extern "C" int ta_fused_update_linf(const float* g, const float* m, float* m_out, const float* delta, float* delta_out,
                                    const float* data, float* xadv_out, const float* scale, float* scale_out, int mean_mode,
                                    float decay, float alpha, float eps, float lo, float hi, int B, int64_t n,
                                    ta_stream_t stream) {
  ta_fused_tail_args a = {};
  a.g = g; a.m = m; a.m_out = m_out; a.delta = delta; a.delta_out = delta_out; a.data = data; a.xadv_out = xadv_out;
  a.scale = scale; a.scale_out = scale_out; a.mean_mode = mean_mode; a.decay = decay; a.alpha = alpha; a.eps = eps; a.lo = lo; a.hi = hi;
  a.B = B; a.n = n;
  return fused_tail_impl(a, nullptr, stream);
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
  NormFold nf;
  const int rc = make_normfold("ta_fused_update_linf_nf", mean_host, std_host, C, plane, n, 1, grad_wrt_xn, &nf);
  if (rc != TA_OK) return rc;
  ta_fused_tail_args a = {};
  a.g = g; a.m = m; a.m_out = m_out; a.delta = delta; a.delta_out = delta_out; a.data = data; a.xadv_out = xn_out;
  a.scale = scale; a.scale_out = scale_out; a.mean_mode = mean_mode; a.decay = decay; a.alpha = alpha; a.eps = eps; a.lo = lo; a.hi = hi;
  a.B = B; a.n = n;
  return fused_tail_impl(a, &nf, stream);
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
  FusedParams p = {};
  p.m = m; p.m_out = m_out; p.delta = delta; p.delta_out = delta_out; p.data = data; p.scale = scale; p.scale_out = scale_out;
  p.decay = decay; p.alpha = alpha; p.eps = eps; p.lo = lo; p.hi = hi; p.n = n;
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
