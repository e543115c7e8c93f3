// aten_mean.cu — host policy of the TA_MEAN_TORCH reduction (see aten_mean.cuh) and the standalone kernel behind
// ta_abs_mean_per_sample(mode = TA_MEAN_TORCH): mean|g| per sample with the bits of torch's CUDA
// `grad.abs().mean(dim=(1,2,3))` (transferattack/attack.py:128), for the public get_momentum hook.
#include "aten_mean.cuh"

namespace ta {

static int last_pow2(int64_t n) { int p = 1; while ((int64_t)p * 2 <= n) p *= 2; return p; }
static int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// PyTorch ATen/native/cuda/Reduce.cuh:1033-1178 setReduceConfig<float, float, vt0 = 4, input_vec_size = 4> for a contiguous
// [B, n] fp32 iterator reduced over its stride-1 dimension, "vectorize along input" (restated; oracle/aten_reduce.py).
bool aten_mean_policy(int B, int64_t n, int sm_count, int max_threads_per_sm, int* bw_, int* bh_, int* cpo_) {
  const int kMax = 512;
  if (B < 1 || n < 128 || n % 4 != 0 || sm_count <= 0 || max_threads_per_sm < kMax) return false;
  const int64_t dim0 = n / 4;
  const int d0p = dim0 < kMax ? last_pow2(dim0) : kMax;
  const int d1p = B < kMax ? last_pow2(B) : kMax;
  int bw = d0p < 32 ? d0p : 32;
  int bh = d1p < kMax / bw ? d1p : kMax / bw;
  bw = d0p < kMax / bh ? d0p : kMax / bh;
  if (bw < 32 || bh > 16) return false;
  const int nt = bw * bh;
  int64_t step = bw;
  int64_t vpt = div_up(n, step);                                   // num_inputs counts elements, the steps count vectors (as in ATen)
  const int64_t thr = (int64_t)bh * 16 < 256 ? (int64_t)bh * 16 : 256;
  if (vpt < thr) return false;                                     // warp rows own separate outputs: not restated
  step *= bh;
  vpt = div_up(n, step);
  const int64_t target = (int64_t)sm_count * (max_threads_per_sm / nt);
  int64_t cpo = 1;
  if (vpt >= 256 && B <= target) {
    const int64_t c1 = div_up(target, B), c2 = div_up(vpt, 16), c3 = div_up(vpt, 256);
    const int64_t mn = c1 < c2 ? c1 : c2;
    cpo = mn > c3 ? mn : c3;
  }
  if (cpo > bw) return false;                                      // final tree here: the partials fill one block row
  *bw_ = bw; *bh_ = bh; *cpo_ = (int)cpo;
  return true;
}

static int g_max_threads_per_sm[64];

int aten_mean_plan(const char* who, int B, int64_t n, int cl, AtenMeanCfg* cfg) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (g_max_threads_per_sm[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxThreadsPerMultiProcessor, dev) != cudaSuccess || v <= 0) v = 2048;
    g_max_threads_per_sm[dev] = v;
  }
  int bw, bh, cpo;
  if (!aten_mean_policy(B, n, sm_count(), g_max_threads_per_sm[dev], &bw, &bh, &cpo)) {
    set_error("%s: TA_MEAN_TORCH does not cover B=%d n=%lld (outside the replayed ATen launch family)", who, B, (long long)n);
    return TA_EUNSUPPORTED;
  }
  const int S = bw * bh * cpo;
  if (cl == 0) {                                   // no cluster mapping wanted (column sums in global memory)
    cfg->bw = bw; cfg->bh = bh; cfg->cpo = cpo; cfg->nt = bw * bh; cfg->S = S; cfg->W4 = S;
    cfg->w4_magic = (1ull << 32) / (unsigned long long)S + 1ull;
    cfg->factor = (float)B / (float)((int64_t)B * n);
    return TA_OK;
  }
  if (cl < 1 || S % cl != 0 || S / cl > kAtenMaxW) {
    set_error("%s: TA_MEAN_TORCH: cluster %d does not divide the %d virtual threads into <= %d columns", who, cl, S, kAtenMaxW);
    return TA_EUNSUPPORTED;
  }
  cfg->bw = bw; cfg->bh = bh; cfg->cpo = cpo; cfg->nt = bw * bh; cfg->S = S; cfg->W4 = S / cl;
  cfg->w4_magic = (1ull << 32) / (unsigned long long)cfg->W4 + 1ull;
  cfg->factor = (float)B / (float)((int64_t)B * n);
  return TA_OK;
}

namespace {

// grid = (cluster, B)
// 4-CTA clusters: 4 x B CTAs of 512 threads — one wave at 2 CTAs/SM for B = 64 (an 8-CTA cluster would need 4 CTAs/SM, i.e. a
// 32-register budget that cannot hold a batch of loads)
template <bool PRE>
__global__ void __launch_bounds__(kAtenThreads, 2) aten_abs_mean_kernel(const float* __restrict__ g, float* __restrict__ mean_out,
                                                                     int64_t n, AtenMeanCfg c, MeanPre pre) {
  __shared__ float s_val[kAtenMaxW];
  __shared__ float s_row[kAtenThreads];
  __shared__ float s_blk[kAtenThreads];
  const float4* gp = reinterpret_cast<const float4*>(g + (int64_t)blockIdx.y * n);
  const int64_t nvec = n >> 2;
  const int64_t col0 = (int64_t)cluster_ctarank() * c.W4;
  const float4* ap = (PRE && pre.addend) ? reinterpret_cast<const float4*>(pre.addend + (int64_t)blockIdx.y * n) : nullptr;
  const int pv = (int)pre.plane_vec;                                 // vectors per channel plane (n < 2^31 here)
  for (int col = threadIdx.x; col < c.W4; col += kAtenThreads) {
    // the column's vectors are v0, v0 + S, v0 + 2S, ...: all loads of a batch of 8 rows are issued before the first add (a
    // thread has no other work to hide a DRAM latency per row behind); the adds then run in row order
    ColAcc A;
    const int v0 = (int)col0 + col;
    const int rows = v0 < (int)nvec ? (int)((nvec - v0 + c.S - 1) / c.S) : 0;
    constexpr int NB = PRE ? 4 : 8;                                  // rows per batch of loads in flight
    for (int j0 = 0; j0 < rows; j0 += NB) {
      float4 x[NB], y[PRE ? NB : 1];
#pragma unroll
      for (int u = 0; u < NB; ++u)
        if (j0 + u < rows) {
          x[u] = __ldg(gp + v0 + (int64_t)(j0 + u) * c.S);
          if (PRE && ap) y[u] = __ldg(ap + v0 + (int64_t)(j0 + u) * c.S);
        }
#pragma unroll
      for (int u = 0; u < NB; ++u)
        if (j0 + u < rows) {
          float4 t = x[u];
          if (PRE) {
            if (pv > 0) {
              const int v = v0 + (j0 + u) * c.S;
              t = div4(t, pick4(pre.std, (v >= pv ? 1 : 0) + (v >= 2 * pv ? 1 : 0) + (v >= 3 * pv ? 1 : 0)));
            }
            if (ap) t = add4(t, y[u]);
          }
          aten_column_add(A, t);
        }
    }
    s_val[col] = aten_column_value(A);
  }
  cluster_sync_all();
  const float mu = aten_tree_mean(c, s_val, s_row, s_blk);
  if (cluster_ctarank() == 0 && threadIdx.x == 0) mean_out[blockIdx.y] = mu;
  cluster_sync_all();                         // s_val must outlive every remote read
}

// Normalize's adjoint gin = gout / std[c] (the bits of ta_normalize_bwd) with the thread <-> data mapping of ATen's mean
// reduction over gin: CTA (x, b) owns virtual threads [512 x, 512 x + 512) of sample b, thread t the 128-bit vectors t, t + S, ...
// — so that besides storing gin it can leave that virtual thread's column value of |gin| in col_sums[b * S + t].
// FINISH: ATen's own structure — every CTA reduces its virtual block (block_x_reduce, block_y_reduce) from shared memory and
// stores one partial; the last CTA of a sample to arrive (ticket counter per sample, reset by that CTA) runs global_reduce's
// final tree over the cpo partials → mean_out[b]; no separate launch for the mean, no column values through global memory.
template <bool FINISH>
__global__ void __launch_bounds__(kAtenThreads, 2) normalize_bwd_colsum_kernel(const float* __restrict__ gout, const float* __restrict__ std,
                                                                            float* __restrict__ gin, float* __restrict__ col_sums,
                                                                            float* __restrict__ mean_out, int* __restrict__ counters,
                                                                            int64_t n, AtenMeanCfg cfg, int plane_vec, int C) {
  extern __shared__ __align__(16) float s_cols[];
  __shared__ float s_row[FINISH ? kAtenThreads : 1];
  __shared__ float s_blk[FINISH ? kAtenThreads : 1];
  __shared__ int s_last;
  const int S = cfg.S;
  const int b = blockIdx.y;
  const float4* gp = reinterpret_cast<const float4*>(gout + (int64_t)b * n);
  float4* ip = reinterpret_cast<float4*>(gin + (int64_t)b * n);
  const int nvec = (int)(n >> 2);
  const int col = blockIdx.x * kAtenThreads + threadIdx.x;
  if (col < S) {
    float sd[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) sd[c] = c < C ? __ldg(std + c) : 1.0f;
    const int rows = col < nvec ? (nvec - col + S - 1) / S : 0;
    ColAcc A;
    constexpr int NB = 4;
    for (int j0 = 0; j0 < rows; j0 += NB) {
      float4 x[NB];
#pragma unroll
      for (int u = 0; u < NB; ++u)
        if (j0 + u < rows) x[u] = __ldg(gp + col + (int64_t)(j0 + u) * S);
#pragma unroll
      for (int u = 0; u < NB; ++u)
        if (j0 + u < rows) {
          const int v = col + (j0 + u) * S;
          const float4 t = div4(x[u], pick4(sd, (v >= plane_vec ? 1 : 0) + (v >= 2 * plane_vec ? 1 : 0) + (v >= 3 * plane_vec ? 1 : 0)));
          ip[v] = t;
          aten_column_add(A, t);
        }
    }
    if (FINISH) s_cols[threadIdx.x] = aten_column_value(A);
    else col_sums[(int64_t)b * S + col] = aten_column_value(A);
  } else if (FINISH) {
    s_cols[threadIdx.x] = 0.0f;
  }
  if (FINISH) {
    // CTA x IS ATen's virtual block x of this sample (S = 512 * cpo, thread id = virtual thread id inside the block): its
    // block_x_reduce and block_y_reduce run here, on the values still in this CTA; only the per-block partial goes through global
    // memory (ATen's staging buffer: col_sums[b * S + x], x < cpo) and the last CTA of the sample to arrive runs global_reduce's
    // final tree over the cpo partials.
    __syncthreads();
    {
      AtenMeanCfg one = cfg; one.cpo = 1;                                         // the trees of ONE block over s_cols[0 .. 512)
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const int K = cfg.bw >> 5;
      switch (K) {
        case 1: aten_rows_x_tree<1>(one, ColSrcShared{s_cols}, s_row, warp, lane); break;
        case 2: aten_rows_x_tree<2>(one, ColSrcShared{s_cols}, s_row, warp, lane); break;
        case 4: aten_rows_x_tree<4>(one, ColSrcShared{s_cols}, s_row, warp, lane); break;
        case 8: aten_rows_x_tree<8>(one, ColSrcShared{s_cols}, s_row, warp, lane); break;
        default: aten_rows_x_tree<16>(one, ColSrcShared{s_cols}, s_row, warp, lane); break;
      }
    }
    __syncthreads();
    float* partials = col_sums + (int64_t)b * S;                                  // the first cpo floats of the sample's slice (no column values are stored in this form)
    if (threadIdx.x == 0) {
      float a[16];
#pragma unroll
      for (int y = 0; y < 16; ++y) a[y] = (y < cfg.bh) ? s_row[y] : 0.0f;
      // block_y_reduce, offsets bh/2 .. 1 (bh is a power of two <= 16; levels above bh do not exist)
#pragma unroll
      for (int h = 8; h >= 1; h >>= 1)
        if (h < cfg.bh) {
#pragma unroll
          for (int y = 0; y < h; ++y) a[y] = add_rn(a[y], a[y + h]);
        }
      s_blk[0] = a[0];
    }
    if (threadIdx.x == 0) {
      __stcg(partials + blockIdx.x, s_blk[0]);
      __threadfence();
      s_last = (atomicAdd(counters + b, 1) == (int)gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      const float* pp = partials;
      const int lane = threadIdx.x & 31;
      if (threadIdx.x < 32) {
        float v;
        if (cfg.cpo == 1) {
          v = __ldcg(pp);
        } else if (cfg.cpo <= 32) {
          float a1[1] = {lane < cfg.cpo ? __ldcg(pp + lane) : 0.0f};
          v = aten_x_tree<1>(a1);
        } else {
          float a16[16];
          const int K = cfg.bw >> 5;
#pragma unroll
          for (int k = 0; k < 16; ++k) { const int i = lane + 32 * k; a16[k] = (k < K && i < cfg.cpo) ? __ldcg(pp + i) : 0.0f; }
          v = aten_x_tree<16>(a16);
        }
        if (lane == 0) { mean_out[b] = mul_rn(v, cfg.factor); counters[b] = 0; }
      }
    }
  }
}

// STAGE: the S values are first copied into shared memory with one batch of independent 128-bit loads per thread (one global
// latency instead of one per pair of block rows), then the trees read them from there
template <bool STAGE>
__global__ void __launch_bounds__(kAtenThreads) aten_colsum_tree_kernel(const float* __restrict__ col_sums, float* __restrict__ mean_out,
                                                                       AtenMeanCfg c) {
  extern __shared__ __align__(16) float s_cols[];
  __shared__ float s_row[kAtenThreads];
  __shared__ float s_blk[kAtenThreads];
  const float* src = col_sums + (int64_t)blockIdx.x * c.S;
  float mu;
  if (STAGE) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(s_cols);
    for (int i = threadIdx.x; i < (c.S >> 2); i += kAtenThreads) d4[i] = __ldg(s4 + i);
    __syncthreads();
    mu = aten_tree_mean_src(c, ColSrcShared{s_cols}, s_row, s_blk);
  } else {
    mu = aten_tree_mean_src(c, ColSrcGlobal{src}, s_row, s_blk);
  }
  if (threadIdx.x == 0) mean_out[blockIdx.x] = mu;
}

}  // namespace

int aten_colsum_normalize_bwd(const float* gout, const float* std, float* gin, float* col_sums, float* mean_out, int* counters, int B, int C,
                              int64_t plane, cudaStream_t s) {
  const int64_t n = (int64_t)C * plane;
  if (C < 1 || C > 4 || plane % 4 != 0 || !aligned16(gout) || !aligned16(gin) || !aligned16(col_sums) || n >= ((int64_t)1 << 31) || B > 65535) {
    set_error("ta_normalize_bwd_colsum: needs C <= 4, H*W %% 4 == 0, 16-byte aligned tensors, B <= 65535");
    return TA_EUNSUPPORTED;
  }
  AtenMeanCfg c;
  const int rc = aten_mean_plan("ta_normalize_bwd_colsum", B, n, 0, &c);
  if (rc != TA_OK) return rc;
  dim3 grid((unsigned)((c.S + kAtenThreads - 1) / kAtenThreads), (unsigned)B);
  if (mean_out && counters && c.nt == kAtenThreads && c.S == kAtenThreads * c.cpo) {
    normalize_bwd_colsum_kernel<true><<<grid, kAtenThreads, sizeof(float) * kAtenThreads, s>>>(gout, std, gin, col_sums, mean_out, counters, n, c,
                                                                                            (int)(plane / 4), C);
  } else if (mean_out) {
    set_error("ta_normalize_bwd_colsum: the in-kernel finish needs ATen blocks of %d threads (here %d)", kAtenThreads, c.nt);
    return TA_EUNSUPPORTED;
  } else
    normalize_bwd_colsum_kernel<false><<<grid, kAtenThreads, 0, s>>>(gout, std, gin, col_sums, nullptr, nullptr, n, c, (int)(plane / 4), C);
  count_launch();
  return check_launch("ta_normalize_bwd_colsum");
}

int aten_colsum_tree(const float* col_sums, float* mean_out, int B, int64_t n, cudaStream_t s) {
  AtenMeanCfg c;
  const int rc = aten_mean_plan("ta_abs_mean_from_colsums", B, n, 0, &c);
  if (rc != TA_OK) return rc;
  if (c.cpo * c.bh > kAtenThreads) { set_error("ta_abs_mean_from_colsums: %d block rows exceed the tree kernel's %d", c.cpo * c.bh, kAtenThreads); return TA_EUNSUPPORTED; }
  const size_t smem = sizeof(float) * (size_t)c.S;
  if (smem <= 48 * 1024 && c.S % 4 == 0 && aligned16(col_sums) && tune_get("reduce.tree_stage", 1) != 0)
    aten_colsum_tree_kernel<true><<<(unsigned)B, kAtenThreads, smem, s>>>(col_sums, mean_out, c);
  else
    aten_colsum_tree_kernel<false><<<(unsigned)B, kAtenThreads, 0, s>>>(col_sums, mean_out, c);
  count_launch();
  return check_launch("ta_abs_mean_from_colsums");
}

int aten_abs_mean_launch(const float* g, float* mean_out, int B, int64_t n, const MeanPre* pre, cudaStream_t s) {
  if (!aligned16(g) || (pre && !aligned16(pre->addend))) {
    set_error("ta_abs_mean_per_sample: TA_MEAN_TORCH needs 16-byte aligned rows");
    return TA_EUNSUPPORTED;
  }
  if (n >= ((int64_t)1 << 31)) { set_error("ta_abs_mean_per_sample: TA_MEAN_TORCH serves samples below 2^31 elements"); return TA_EUNSUPPORTED; }
  int cl = tune_get("reduce.cluster", 0);
  if (cl <= 0) cl = 4;
  AtenMeanCfg c;
  int rc = aten_mean_plan("ta_abs_mean_per_sample", B, n, cl, &c);
  if (rc != TA_OK && cl < 8 && tune_get("reduce.cluster", 0) <= 0) rc = aten_mean_plan("ta_abs_mean_per_sample", B, n, 8, &c), cl = 8;   // more columns than s_val holds
  if (rc != TA_OK) return rc;
  if (pre && (pre->addend || pre->plane_vec > 0))
    return launch_cluster("ta_abs_mean_per_sample[torch order]", aten_abs_mean_kernel<true>, cl, B, kAtenThreads, 0, s, g, mean_out, n, c, *pre);
  return launch_cluster("ta_abs_mean_per_sample[torch order]", aten_abs_mean_kernel<false>, cl, B, kAtenThreads, 0, s, g, mean_out, n, c,
                        MeanPre{});
}

}  // namespace ta

// ATen's launch policy for x.mean over the last dimension of a contiguous [B, n] fp32 tensor on a device with `sm_count` SMs and
// `max_threads_per_sm` resident threads per SM (host-only: callable without a GPU; tests compare it with oracle/aten_reduce.py).
extern "C" int ta_aten_mean_policy(int B, int64_t n, int sm_count, int max_threads_per_sm, int* block_w, int* block_h,
                                   int* ctas_per_output) {
  int bw = 0, bh = 0, cpo = 0;
  if (!ta::aten_mean_policy(B, n, sm_count, max_threads_per_sm, &bw, &bh, &cpo)) {
    ta::set_error("ta_aten_mean_policy: B=%d n=%lld is outside the replayed launch family", B, (long long)n);
    return TA_EUNSUPPORTED;
  }
  if (block_w) *block_w = bw;
  if (block_h) *block_h = bh;
  if (ctas_per_output) *ctas_per_output = cpo;
  return TA_OK;
}

// Normalize's adjoint gin = gout / std[c] (utils.py:72-79; same bits as ta_normalize_bwd) that ALSO leaves, per sample, the S
// column values of |gin| of torch's `gin.abs().mean(dim=(1,2,3))` reduction (attack.py:128) in col_sums [B, S], S = block_w *
// block_h * ctas_per_output of ta_aten_mean_policy: ta_abs_mean_from_colsums then finishes that mean (bit-identical to torch's)
// from 4*S bytes per sample instead of a pass over the gradient — or, with mean_out [B] and counters [B] (int, zero before the
// first call, left zero), the last CTA of every sample finishes it inside this launch. TA_EUNSUPPORTED outside the replayed family.
extern "C" int ta_normalize_bwd_colsum(const float* gout, const float* std, float* gin, float* col_sums, float* mean_out, int* counters,
                                       int B, int C, int64_t plane, ta_stream_t stream) {
  TA_REQUIRE(gout && std && gin && col_sums && B > 0 && C > 0 && plane > 0, "ta_normalize_bwd_colsum: bad arguments");
  TA_REQUIRE((mean_out == nullptr) == (counters == nullptr), "ta_normalize_bwd_colsum: mean_out and counters go together");
  return ta::aten_colsum_normalize_bwd(gout, std, gin, col_sums, mean_out, counters, B, C, plane, (cudaStream_t)stream);
}
extern "C" int ta_abs_mean_from_colsums(const float* col_sums, float* mean_out, int B, int64_t n, ta_stream_t stream) {
  TA_REQUIRE(col_sums && mean_out && B > 0 && n > 0, "ta_abs_mean_from_colsums: bad arguments");
  return ta::aten_colsum_tree(col_sums, mean_out, B, n, (cudaStream_t)stream);
}
