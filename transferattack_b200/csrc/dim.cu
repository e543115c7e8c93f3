// dim.cu — DIM's resize → zero-pad → resize (input_transformation/dim.py:42-68) as ONE gather kernel, and its
// exact adjoint in deterministic gather form (ATen's bilinear backward scatters with atomicAdd).
//
// Geometry (one (rnd, pad_top, pad_left) per batch, as in the reference):
//   y1 = bilinear(x: S x S -> rnd x rnd)     taps1[q]: source rows/cols of y1 index q
//   y2 = zero-pad(y1) to R x R at (top, left)
//   out = bilinear(y2: R x R -> S x S)       taps2[o]: y2 rows/cols of output index o
// ATen index math (align_corners=False): scale = (float)in/(float)out; src = max(0, fmaf(scale, dst+0.5f, -0.5f));
// i0 = (int)src; i1 = i0 + (i0 < in-1); l1 = src - i0; l0 = 1 - l1;
// val = hl0*(wl0*p00 + wl1*p01) + hl1*(wl0*p10 + wl1*p11), every product/sum rounded once (no FMA).
//
// Each CTA owns a band of RB (= 16) output rows of one plane. The source rows that band depends on are contiguous in
// memory, so they are staged into shared memory with one bulk-TMA copy (cp.async.bulk + mbarrier); the y1 band is
// formed once in shared memory (not 4x per output) and the outputs gather from it. HBM traffic: 4 B/elem in (+ halo
// rows) and 4 B/elem out.
#include "common.cuh"

using namespace ta;

namespace {

constexpr int RB = 16;         // output rows per CTA
constexpr int kThreads = 256;

struct Tap { int i0, i1; float l0, l1; };

__device__ __forceinline__ Tap make_tap(int in, float scale, int d) {
  float src = fmaf(scale, (float)d + 0.5f, -0.5f);      // the reference's single-rounding index (ATen area_pixel_compute_source_index)
  if (src < 0.0f) src = 0.0f;
  Tap t;
  t.i0 = (int)src;
  t.i1 = t.i0 + ((t.i0 < in - 1) ? 1 : 0);
  t.l1 = sub_rn(src, (float)t.i0);
  t.l0 = sub_rn(1.0f, t.l1);
  return t;
}

// ATen's expression is  hl0*(wl0*p00 + wl1*p01) + hl1*(wl0*p10 + wl1*p11).
// mode 1 (default): top = fma(wl0,p00, wl1*p01), bot likewise, val = fma(hl0,top, hl1*bot) — the contraction nvcc applied
//   to torch's own CUDA kernel; MEASURED on B200 (tools/diag_dim_aten.py, profiles/diag_dim_r1.json): 0 differing bits
//   against F.interpolate -> F.pad -> F.interpolate for every geometry tried, so DIM's forward is bit-identical to the
//   reference's GPU path. mode 0: every product and sum rounded separately (closest to ATen's CPU kernel, used with the
//   CPU goldens). modes 2-4: the other contraction orders (kept for the diagnostic).
__device__ __forceinline__ float blend(int mode, float hl0, float hl1, float wl0, float wl1, float p00, float p01, float p10, float p11) {
  float top, bot;
  if (mode == 0) {
    top = add_rn(mul_rn(wl0, p00), mul_rn(wl1, p01));
    bot = add_rn(mul_rn(wl0, p10), mul_rn(wl1, p11));
    return add_rn(mul_rn(hl0, top), mul_rn(hl1, bot));
  }
  if (mode == 1 || mode == 3) {
    top = fmaf(wl0, p00, mul_rn(wl1, p01));
    bot = fmaf(wl0, p10, mul_rn(wl1, p11));
  } else {
    top = fmaf(wl1, p01, mul_rn(wl0, p00));
    bot = fmaf(wl1, p11, mul_rn(wl0, p10));
  }
  if (mode == 1 || mode == 4) return fmaf(hl0, top, mul_rn(hl1, bot));
  return fmaf(hl1, bot, mul_rn(hl0, top));
}

struct DimGeom { int S, rnd, R, top, left; int y1_rows_max, src_rows_max; int blend; };

// shared-memory carve-up (dynamic): [taps2: S][taps1: rnd][y1 band: y1_rows_max*rnd floats][src band: src_rows_max*S floats]
template <bool TMA_STAGE>
__global__ void __launch_bounds__(kThreads) dim_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, DimGeom gm) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  const int S = gm.S, rnd = gm.rnd, R = gm.R, top = gm.top, left = gm.left;
  // src band first (bulk-TMA destination must be 16-B aligned: offset 0 of the 128-B aligned window)
  float* s_src = reinterpret_cast<float*>(smem_raw);
  const size_t src_bytes = TMA_STAGE ? (((size_t)gm.src_rows_max * S * 4 + 15) & ~(size_t)15) : 0;
  Tap* taps2 = reinterpret_cast<Tap*>(smem_raw + src_bytes);
  Tap* taps1 = taps2 + S;
  float* s_y1 = reinterpret_cast<float*>(taps1 + rnd);

  const int tid = threadIdx.x;
  const int oy0 = blockIdx.x * RB;
  const int oy1 = min(oy0 + RB, S) - 1;                 // inclusive
  const float* xp = x + (int64_t)blockIdx.y * S * S;
  float* op = out + (int64_t)blockIdx.y * S * S;

  const float scale2 = (float)R / (float)S, scale1 = (float)S / (float)rnd;
  for (int i = tid; i < S; i += kThreads) taps2[i] = make_tap(R, scale2, i);
  for (int i = tid; i < rnd; i += kThreads) taps1[i] = make_tap(S, scale1, i);
  if (TMA_STAGE && tid == 0) { mbar_init(&s_bar, 1); mbar_fence_init(); }
  __syncthreads();

  // y1 rows this band touches (may be empty when the band maps entirely into the padding)
  const int pr0 = taps2[oy0].i0, pr1 = taps2[oy1].i1;
  const int q0 = max(pr0 - top, 0), q1 = min(pr1 - top, rnd - 1);
  const bool any = q0 <= q1;
  int sr0 = 0, sr1 = -1;
  if (any) { sr0 = taps1[q0].i0; sr1 = taps1[q1].i1; }

  if (TMA_STAGE && any) {
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)((sr1 - sr0 + 1) * S * 4);
      mbar_expect_tx(&s_bar, bytes);
      tma_bulk_g2s(s_src, xp + (int64_t)sr0 * S, bytes, &s_bar);
    }
    mbar_wait(&s_bar, 0);
  }

  // stage 1: y1 band into shared memory
  if (any) {
    const int rows = q1 - q0 + 1;
    for (int e = tid; e < rows * rnd; e += kThreads) {
      const int q = q0 + e / rnd, qx = e % rnd;
      const Tap th = taps1[q], tw = taps1[qx];
      float p00, p01, p10, p11;
      if (TMA_STAGE) {
        const float* r0 = s_src + (th.i0 - sr0) * S;
        const float* r1 = s_src + (th.i1 - sr0) * S;
        p00 = r0[tw.i0]; p01 = r0[tw.i1]; p10 = r1[tw.i0]; p11 = r1[tw.i1];
      } else {
        const float* r0 = xp + (int64_t)th.i0 * S;
        const float* r1 = xp + (int64_t)th.i1 * S;
        p00 = __ldg(r0 + tw.i0); p01 = __ldg(r0 + tw.i1); p10 = __ldg(r1 + tw.i0); p11 = __ldg(r1 + tw.i1);
      }
      s_y1[e] = blend(gm.blend, th.l0, th.l1, tw.l0, tw.l1, p00, p01, p10, p11);
    }
  }
  __syncthreads();

  // stage 2: outputs gather from the (implicitly zero-padded) y1 band
  const int nout = (oy1 - oy0 + 1) * S;
  for (int e = tid; e < nout; e += kThreads) {
    const int oy = oy0 + e / S, ox = e % S;
    const Tap th = taps2[oy], tw = taps2[ox];
    const int ya = th.i0 - top, yb = th.i1 - top, xa = tw.i0 - left, xb = tw.i1 - left;
    const bool ya_in = any && ya >= q0 && ya <= q1, yb_in = any && yb >= q0 && yb <= q1;
    const bool xa_in = xa >= 0 && xa < rnd, xb_in = xb >= 0 && xb < rnd;
    const float v00 = (ya_in && xa_in) ? s_y1[(ya - q0) * rnd + xa] : 0.0f;
    const float v01 = (ya_in && xb_in) ? s_y1[(ya - q0) * rnd + xb] : 0.0f;
    const float v10 = (yb_in && xa_in) ? s_y1[(yb - q0) * rnd + xa] : 0.0f;
    const float v11 = (yb_in && xb_in) ? s_y1[(yb - q0) * rnd + xb] : 0.0f;
    op[(int64_t)oy * S + ox] = blend(gm.blend, th.l0, th.l1, tw.l0, tw.l1, v00, v01, v10, v11);
  }
}

// ---- adjoint --------------------------------------------------------------------------------------------------------
// weight with which 1-D tap `t` (of some output index) reads input index `i`: l0 if i0 == i, plus l1 if i1 == i
__device__ __forceinline__ float tap_w(const Tap& t, int i) {
  float w = 0.0f;
  if (t.i0 == i) w = t.l0;
  if (t.i1 == i) w = add_rn(w, t.l1);
  return w;
}

// smem: [taps2: S][taps1: rnd][inv2 lo/hi: 2R ints][inv1 lo/hi: 2S ints][g1 band: g1_rows_max * rnd floats]
__global__ void __launch_bounds__(kThreads) dim_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, DimGeom gm) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = gm.S, rnd = gm.rnd, R = gm.R, top = gm.top, left = gm.left;
  Tap* taps2 = reinterpret_cast<Tap*>(smem_raw);
  Tap* taps1 = taps2 + S;
  int* inv2_lo = reinterpret_cast<int*>(taps1 + rnd);
  int* inv2_hi = inv2_lo + R;
  int* inv1_lo = inv2_hi + R;
  int* inv1_hi = inv1_lo + S;
  float* s_g1 = reinterpret_cast<float*>(inv1_hi + S);

  const int tid = threadIdx.x;
  const int sy0 = blockIdx.x * RB;
  const int sy1 = min(sy0 + RB, S) - 1;
  const float* gp = gout + (int64_t)blockIdx.y * S * S;
  float* ip = gin + (int64_t)blockIdx.y * S * S;

  const float scale2 = (float)R / (float)S, scale1 = (float)S / (float)rnd;
  for (int i = tid; i < S; i += kThreads) { taps2[i] = make_tap(R, scale2, i); inv1_lo[i] = 0x7fffffff; inv1_hi[i] = -1; }
  for (int i = tid; i < rnd; i += kThreads) taps1[i] = make_tap(S, scale1, i);
  for (int i = tid; i < R; i += kThreads) { inv2_lo[i] = 0x7fffffff; inv2_hi[i] = -1; }
  __syncthreads();
  // inverse ranges (min/max → order-independent): which outputs o read y2 index p; which y1 indices q read source s
  for (int o = tid; o < S; o += kThreads) {
    atomicMin(&inv2_lo[taps2[o].i0], o); atomicMax(&inv2_hi[taps2[o].i1], o);
    atomicMin(&inv2_lo[taps2[o].i1], o); atomicMax(&inv2_hi[taps2[o].i0], o);
  }
  for (int q = tid; q < rnd; q += kThreads) {
    atomicMin(&inv1_lo[taps1[q].i0], q); atomicMax(&inv1_hi[taps1[q].i1], q);
    atomicMin(&inv1_lo[taps1[q].i1], q); atomicMax(&inv1_hi[taps1[q].i0], q);
  }
  __syncthreads();

  // y1 rows feeding this band of source rows (taps are monotone, so the union of ranges is a range)
  int q0 = 0x7fffffff, q1 = -1;
  for (int sy = sy0; sy <= sy1; ++sy) { q0 = min(q0, inv1_lo[sy]); q1 = max(q1, inv1_hi[sy]); }
  const bool any = q0 <= q1;

  // stage 1 (adjoint of the second resize, cropped to the pad window): g1[q][qx] = g2[q+top][qx+left]
  if (any) {
    const int rows = q1 - q0 + 1;
    for (int e = tid; e < rows * rnd; e += kThreads) {
      const int py = q0 + e / rnd + top, px = e % rnd + left;
      float acc = 0.0f;
      for (int oy = inv2_lo[py]; oy <= inv2_hi[py]; ++oy) {
        const float wy = tap_w(taps2[oy], py);
        if (wy == 0.0f) continue;
        for (int ox = inv2_lo[px]; ox <= inv2_hi[px]; ++ox) {
          const float wx = tap_w(taps2[ox], px);
          if (wx == 0.0f) continue;
          acc = add_rn(acc, mul_rn(mul_rn(wy, wx), __ldg(gp + (int64_t)oy * S + ox)));
        }
      }
      s_g1[e] = acc;
    }
  }
  __syncthreads();

  // stage 2 (adjoint of the first resize)
  const int nout = (sy1 - sy0 + 1) * S;
  for (int e = tid; e < nout; e += kThreads) {
    const int sy = sy0 + e / S, sx = e % S;
    float acc = 0.0f;
    for (int q = inv1_lo[sy]; q <= inv1_hi[sy]; ++q) {
      const float wy = tap_w(taps1[q], sy);
      if (wy == 0.0f) continue;
      for (int qx = inv1_lo[sx]; qx <= inv1_hi[sx]; ++qx) {
        const float wx = tap_w(taps1[qx], sx);
        if (wx == 0.0f) continue;
        acc = add_rn(acc, mul_rn(mul_rn(wy, wx), s_g1[(q - q0) * rnd + qx]));
      }
    }
    ip[(int64_t)sy * S + sx] = acc;
  }
}

int check_geom(const char* who, int planes, int S, int rnd, int R, int top, int left) {
  TA_REQUIRE(planes > 0 && S > 0, "%s: empty shape", who);
  TA_REQUIRE(rnd >= 1 && R >= rnd && top >= 0 && left >= 0 && top + rnd <= R && left + rnd <= R,
             "%s: bad geometry S=%d rnd=%d R=%d top=%d left=%d", who, S, rnd, R, top, left);
  TA_REQUIRE(planes <= 65535, "%s: planes=%d exceeds 65535", who, planes);
  return TA_OK;
}

// conservative row bounds for the shared-memory bands
int band_rows(int rows_out, int in, int out) {   // input rows touched by `rows_out` consecutive output rows of an in->out resize
  return (int)((double)rows_out * (double)in / (double)out) + 3;
}

}  // namespace

extern "C" {

int ta_dim_fwd(const float* x, float* out, int planes, int S, int rnd, int R, int pad_top, int pad_left, ta_stream_t stream) {
  TA_REQUIRE(x && out, "ta_dim_fwd: null pointer");
  int rc = check_geom("ta_dim_fwd", planes, S, rnd, R, pad_top, pad_left);
  if (rc != TA_OK) return rc;
  DimGeom gm{S, rnd, R, pad_top, pad_left, 0, 0, tune_get("dim.blend", 1)};
  gm.y1_rows_max = band_rows(RB, R, S);
  if (gm.y1_rows_max > rnd) gm.y1_rows_max = rnd;
  gm.src_rows_max = band_rows(gm.y1_rows_max, S, rnd);
  if (gm.src_rows_max > S) gm.src_rows_max = S;
  const bool can_tma = (S % 4 == 0) && aligned16(x) && tune_get("dim.tma", 1) != 0;
  const size_t src_bytes = can_tma ? (((size_t)gm.src_rows_max * S * 4 + 15) & ~(size_t)15) : 0;
  const size_t smem = src_bytes + sizeof(Tap) * (size_t)(S + rnd) + sizeof(float) * (size_t)gm.y1_rows_max * rnd;
  TA_REQUIRE(smem <= 200 * 1024, "ta_dim_fwd: image size S=%d needs %zu B of shared memory per CTA", S, smem);
  auto k = can_tma ? dim_fwd_kernel<true> : dim_fwd_kernel<false>;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("ta_dim_fwd: smem attribute: %s", cudaGetErrorString(e)); cudaGetLastError(); return TA_ECUDA; }
  }
  dim3 grid((unsigned)((S + RB - 1) / RB), (unsigned)planes);
  k<<<grid, kThreads, smem, (cudaStream_t)stream>>>(x, out, gm);
  count_launch();
  return check_launch("ta_dim_fwd");
}

int ta_dim_bwd(const float* gout, float* gin, int planes, int S, int rnd, int R, int pad_top, int pad_left, ta_stream_t stream) {
  TA_REQUIRE(gout && gin, "ta_dim_bwd: null pointer");
  int rc = check_geom("ta_dim_bwd", planes, S, rnd, R, pad_top, pad_left);
  if (rc != TA_OK) return rc;
  DimGeom gm{S, rnd, R, pad_top, pad_left, 0, 0, 0};
  // y1 rows reading RB consecutive source rows of the S -> rnd resize
  gm.y1_rows_max = (int)((double)(RB + 1) * (double)rnd / (double)S) + 3;
  if (gm.y1_rows_max > rnd) gm.y1_rows_max = rnd;
  const size_t smem = sizeof(Tap) * (size_t)(S + rnd) + sizeof(int) * (size_t)(2 * R + 2 * S) +
                      sizeof(float) * (size_t)gm.y1_rows_max * rnd;
  TA_REQUIRE(smem <= 200 * 1024, "ta_dim_bwd: image size S=%d needs %zu B of shared memory per CTA", S, smem);
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(dim_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("ta_dim_bwd: smem attribute: %s", cudaGetErrorString(e)); cudaGetLastError(); return TA_ECUDA; }
  }
  dim3 grid((unsigned)((S + RB - 1) / RB), (unsigned)planes);
  dim_bwd_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(gout, gin, gm);
  count_launch();
  return check_launch("ta_dim_bwd");
}

}  // extern "C"
