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

}  // namespace

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
