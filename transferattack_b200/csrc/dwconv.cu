// dwconv.cu — TIM's depthwise convolution of the input gradient (input_transformation/tim.py:68-73):
//   out = conv2d(g, K[C,1,ks,ks], stride 1, zero 'same' padding, groups=C)   (cross-correlation)
//
// ta_dwconv2d_sep: the kernels tim.py:42-66 generates (gaussian / uniform / linear) are rank-1, K = outer(kcol, krow);
//   the convolution is done as a row pass then a column pass inside one CTA (intermediate in shared memory), 2*ks
//   FMAs per output instead of ks*ks, which moves the op from FFMA-bound (225 MAC/elem at ks=15) back to HBM-bound
//   (8 B/elem). Accumulation: fp32 FMA chains from 0 in tap order j = 0..ks-1 then i = 0..ks-1 — the order the
//   oracle (orc_dwconv2d_sep) replays, so kernel and oracle agree bit for bit.
// ta_dwconv2d: any [C,ks,ks] kernel, fp32 FMA chain in (ky,kx) raster order (orc_dwconv2d order).
//
// Tiling: 32x32 outputs per CTA, 256 threads, halo tile (32+ks-1)^2 in shared memory with an odd row stride
// (bank-conflict-free 4-wide register blocking); zero padding comes from the guarded tile load.
#include "common.cuh"

#include <string.h>

using namespace ta;

namespace {

constexpr int TH = 32, TW = 32, kThreads = 256, kMaxKs = 31;

__device__ __forceinline__ int odd_up(int v) { return v | 1; }

// cooperative guarded load of the halo tile: s_in[(TH+ks-1)][IS], zero outside the image
__device__ __forceinline__ void load_tile(const float* __restrict__ gp, float* s_in, int IS, int ks, int H, int W, int y0, int x0) {
  const int r = ks >> 1, th = TH + ks - 1, tw = TW + ks - 1;
  for (int e = threadIdx.x; e < th * tw; e += kThreads) {
    const int ty = e / tw, tx = e % tw;
    const int yy = y0 + ty - r, xx = x0 + tx - r;
    s_in[ty * IS + tx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(gp + (int64_t)yy * W + xx) : 0.0f;
  }
}

// KS > 0: compile-time size, weights in registers, 4-wide register blocking. KS == 0: runtime size, generic loops.
template <int KS>
__global__ void __launch_bounds__(kThreads) dwconv_sep_kernel(const float* __restrict__ g, const float* __restrict__ kcol,
                                                              const float* __restrict__ krow, int ks_rt, float* __restrict__ out,
                                                              int C, int H, int W) {
  extern __shared__ __align__(16) float smem[];
  const int ks = KS > 0 ? KS : ks_rt;
  const int th = TH + ks - 1;
  const int IS = odd_up(TW + ks - 1);
  float* s_in = smem;                 // [th][IS]
  float* s_tmp = s_in + th * IS;      // [th][TW]
  float* s_kr = s_tmp + th * TW;      // [ks]
  float* s_kc = s_kr + kMaxKs;        // [ks]

  const int plane = blockIdx.z, c = plane % C;
  const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
  const float* gp = g + (int64_t)plane * H * W;
  float* op = out + (int64_t)plane * H * W;
  const int tid = threadIdx.x;
  if (tid < ks) { s_kr[tid] = __ldg(krow + c * ks + tid); s_kc[tid] = __ldg(kcol + c * ks + tid); }
  load_tile(gp, s_in, IS, ks, H, W, y0, x0);
  __syncthreads();

  if (KS > 0) {
    float wr[KS > 0 ? KS : 1];
#pragma unroll
    for (int j = 0; j < KS; ++j) wr[j] = s_kr[j];
    // row pass: th rows x 8 groups of 4 outputs
    for (int e = tid; e < th * (TW / 4); e += kThreads) {
      const int y = e / (TW / 4), xg = e % (TW / 4);
      const float* row = s_in + y * IS + 4 * xg;
      float v[(KS > 0 ? KS : 1) + 3];
#pragma unroll
      for (int t = 0; t < KS + 3; ++t) v[t] = row[t];
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        a0 = fmaf(wr[j], v[j], a0); a1 = fmaf(wr[j], v[j + 1], a1);
        a2 = fmaf(wr[j], v[j + 2], a2); a3 = fmaf(wr[j], v[j + 3], a3);
      }
      float* t4 = s_tmp + y * TW + 4 * xg;
      t4[0] = a0; t4[1] = a1; t4[2] = a2; t4[3] = a3;
    }
    __syncthreads();
    float wc[KS > 0 ? KS : 1];
#pragma unroll
    for (int i = 0; i < KS; ++i) wc[i] = s_kc[i];
    // column pass: 8 groups of 4 rows x 32 columns = 256 items, one per thread
    {
      const int x = tid % TW, yg = tid / TW;
      float v[(KS > 0 ? KS : 1) + 3];
#pragma unroll
      for (int t = 0; t < KS + 3; ++t) v[t] = s_tmp[(4 * yg + t) * TW + x];
      float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < KS; ++i) {
        a[0] = fmaf(wc[i], v[i], a[0]); a[1] = fmaf(wc[i], v[i + 1], a[1]);
        a[2] = fmaf(wc[i], v[i + 2], a[2]); a[3] = fmaf(wc[i], v[i + 3], a[3]);
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int yy = y0 + 4 * yg + o, xx = x0 + x;
        if (yy < H && xx < W) op[(int64_t)yy * W + xx] = a[o];
      }
    }
  } else {
    for (int e = tid; e < th * TW; e += kThreads) {
      const int y = e / TW, x = e % TW;
      float acc = 0.f;
      for (int j = 0; j < ks; ++j) acc = fmaf(s_kr[j], s_in[y * IS + x + j], acc);
      s_tmp[e] = acc;
    }
    __syncthreads();
    for (int e = tid; e < TH * TW; e += kThreads) {
      const int y = e / TW, x = e % TW;
      float acc = 0.f;
      for (int i = 0; i < ks; ++i) acc = fmaf(s_kc[i], s_tmp[(y + i) * TW + x], acc);
      const int yy = y0 + y, xx = x0 + x;
      if (yy < H && xx < W) op[(int64_t)yy * W + xx] = acc;
    }
  }
}


// ---- band variant of the separable convolution (the TIM hot case: W % 4 == 0, W <= 512) ---------------------------------
// One CTA = BH output rows x the full width of one plane. The (BH + ks - 1) input rows it needs are whole image rows, i.e.
// contiguous in memory: each valid row is brought into shared memory by one bulk-TMA copy (cp.async.bulk, all rows on one
// mbarrier) into a row-padded layout whose left/right margins and out-of-image rows are zero (that IS the 'same' zero
// padding). Row pass: 4 outputs per item from 128-bit conflict-free LDS; column pass: 4 rows per item, stride-1 LDS.
// Same FMA order as orc_dwconv2d_sep (taps ascending from 0) → bit-identical to the tile kernel and the oracle.
constexpr int BH = 32, kBandThreads = 512;

template <int KS>
__global__ void __launch_bounds__(kBandThreads) dwconv_sep_band_kernel(const float* __restrict__ g, const float* __restrict__ kcol,
                                                                       const float* __restrict__ krow, float* __restrict__ out,
                                                                       int C, int H, int W) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  constexpr int R = KS / 2;
  constexpr int PADX = (R + 3) & ~3;               // left/right zero margin, multiple of 4 floats (16-B aligned rows)
  constexpr int NV = (KS + 3 + (PADX - R) + 3) / 4;       // aligned float4 loads covering the 4-output window
  const int WP = W + 2 * PADX;
  const int rows = BH + KS - 1;
  float* s_in = reinterpret_cast<float*>(smem_raw);            // [rows][WP]
  float* s_tmp = s_in + rows * WP;                               // [rows][W]
  const int tid = threadIdx.x;
  const int plane = blockIdx.y, c = plane % C;
  const int y0 = blockIdx.x * BH;
  const float* gp = g + (int64_t)plane * H * W;
  float* op = out + (int64_t)plane * H * W;

  float wr[KS], wc[KS];
#pragma unroll
  for (int j = 0; j < KS; ++j) { wr[j] = __ldg(krow + c * KS + j); wc[j] = __ldg(kcol + c * KS + j); }

  // valid input rows of this band: image rows [ya, yb)
  const int ya = max(y0 - R, 0), yb = min(y0 + BH + R, H);
  if (tid == 0) { mbar_init(&s_bar, 1); mbar_fence_init(); }
  // zero the margins of every row and the rows that fall outside the image (disjoint from the TMA destinations); only the
  // first / last band of a plane has such rows: [0, top_inv) and [rows - bot_inv, rows)
  for (int e = tid; e < rows * 2 * PADX; e += kBandThreads) {
    const int r = e / (2 * PADX), q = e % (2 * PADX);       // constants: shifts / masks
    s_in[r * WP + (q < PADX ? q : W + q)] = 0.0f;
  }
  const int top_inv = ya - (y0 - R), bot_inv = (y0 + BH + R) - yb;
  for (int e = tid; e < (top_inv + bot_inv) * W; e += kBandThreads) {
    const int k = e / W, x = e - k * W;
    const int r = k < top_inv ? k : rows - bot_inv + (k - top_inv);
    s_in[r * WP + PADX + x] = 0.0f;
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s_bar, (uint32_t)((yb - ya) * W * 4));
    for (int yy = ya; yy < yb; ++yy) tma_bulk_g2s(s_in + (yy - (y0 - R)) * WP + PADX, gp + (int64_t)yy * W, (uint32_t)(W * 4), &s_bar);
  }
  mbar_wait(&s_bar, 0);

  // row pass: tmp[r][x] = sum_j krow[j] * in[r][x + j - R]
  const int groups = W >> 2;
  const int dq = kBandThreads / groups, dr = kBandThreads % groups;   // advancing e by the block size without a division
  int r = tid / groups, xg = tid % groups;
  for (int e = tid; e < rows * groups; e += kBandThreads, r += dq, xg += dr) {
    if (xg >= groups) { xg -= groups; ++r; }
    // outputs x = 4xg..4xg+3 read padded columns 4xg + (PADX - R) + [0, KS + 3)
    const float4* row4 = reinterpret_cast<const float4*>(s_in + r * WP + 4 * xg);
    float v[4 * NV];
#pragma unroll
    for (int t = 0; t < NV; ++t) {
      const float4 q = row4[t];
      v[4 * t] = q.x; v[4 * t + 1] = q.y; v[4 * t + 2] = q.z; v[4 * t + 3] = q.w;
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 0; j < KS; ++j) {
      const int b = (PADX - R) + j;
      a0 = fmaf(wr[j], v[b], a0); a1 = fmaf(wr[j], v[b + 1], a1); a2 = fmaf(wr[j], v[b + 2], a2); a3 = fmaf(wr[j], v[b + 3], a3);
    }
    *reinterpret_cast<float4*>(s_tmp + r * W + 4 * xg) = make_float4(a0, a1, a2, a3);
  }
  __syncthreads();

  // column pass: out[y][x] = sum_i kcol[i] * tmp[y + i][x]; item = (4 rows, 1 column)
  const int cq = kBandThreads / W, cr = kBandThreads % W;
  int yg = tid / W, x = tid % W;
  for (int e = tid; e < (BH / 4) * W; e += kBandThreads, yg += cq, x += cr) {
    if (x >= W) { x -= W; ++yg; }
    float v[KS + 3];
#pragma unroll
    for (int t = 0; t < KS + 3; ++t) v[t] = s_tmp[(4 * yg + t) * W + x];
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < KS; ++i) {
      a[0] = fmaf(wc[i], v[i], a[0]); a[1] = fmaf(wc[i], v[i + 1], a[1]); a[2] = fmaf(wc[i], v[i + 2], a[2]); a[3] = fmaf(wc[i], v[i + 3], a[3]);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int yy = y0 + 4 * yg + o;
      if (yy < H) op[(int64_t)yy * W + x] = a[o];
    }
  }
}

template <int KS>
int launch_band(const float* g, const float* kcol, const float* krow, float* out, int B, int C, int H, int W, cudaStream_t s) {
  constexpr int R = KS / 2, PADX = (R + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)(BH + KS - 1) * (W + 2 * PADX) + (size_t)(BH + KS - 1) * W);
  auto k = dwconv_sep_band_kernel<KS>;
  static SmemOptIn optin = {};
  const int rc = ensure_dyn_smem("ta_dwconv2d_sep", k, smem, optin);
  if (rc != TA_OK) return rc;
  dim3 grid((unsigned)((H + BH - 1) / BH), (unsigned)(B * C));
  k<<<grid, kBandThreads, smem, s>>>(g, kcol, krow, out, C, H, W);
  count_launch();
  return check_launch("ta_dwconv2d_sep[band]");
}

// ---- register-sliding variant of the separable convolution (default for the TIM hot case) ---------------------------------
// One CTA = one band of BH output rows x the full width of one plane, ONE thread per 4 adjacent columns. The band's
// BH + KS - 1 input rows are staged exactly like the band kernel (one bulk-TMA copy per image row into a zero-margined
// layout), but in chunks of CH rows with one mbarrier each, so the first rows are consumed while the rest are in flight.
// A thread then walks down the band ONCE: for input row r it forms the row pass of its 4 columns in registers
// (5 x LDS.128, KS x 4 FMA) and immediately scatters that value into the KS output rows it contributes to,
//   acc[(r - i) mod KS] = fma(kcol[i], tmp_r, acc[(r - i) mod KS]),  i = 0..KS-1,
// a rotating file of KS x 4 accumulators whose slot index is static because the row loop is unrolled KS-fold. Output row
// y = r - (KS-1) is complete after input row r and leaves as one 128-bit store. The intermediate never touches shared
// memory, there is no second pass, no per-item index arithmetic, and per output the FMA chain still runs tap 0..KS-1 from
// 0 in both directions → bit-identical to the band / tile kernels and to orc_dwconv2d_sep.
// PW = true: the weights are kernel parameters (constant bank operands of the FMAs, no registers) — used when the host knows
// them (ta_dwconv2d_sep_hw); PW = false: loaded once per thread from the device arrays.
template <int KS> struct SepWeights { float kr[KS]; float kc[KS]; };

template <int KS> struct RsGeom {
  static constexpr int R = KS / 2;
  static constexpr int PADX = (R + 3) & ~3;
  static constexpr int OFF = PADX - R;                    // first padded column read by output column 0
  static constexpr int NV = (OFF + KS + 3 + 3) / 4;       // float4 loads covering the 4-output window
  static constexpr int M = (8 + KS - 1) / KS;             // chunk = M * KS rows (>= 8)
  static constexpr int CH = M * KS;
};

template <int KS, int BHR, bool PW>
__global__ void __launch_bounds__(128) dwconv_sep_rs_kernel(const float* __restrict__ g, const float* __restrict__ kcol,
                                                            const float* __restrict__ krow,
                                                            const __grid_constant__ SepWeights<KS> wp, float* __restrict__ out,
                                                            int C, int H, int W) {
  using G = RsGeom<KS>;
  constexpr int R = G::R, PADX = G::PADX, OFF = G::OFF, NV = G::NV, CH = G::CH;
  constexpr int ROWS = BHR + KS - 1;
  constexpr int NCH = (ROWS + CH - 1) / CH;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[NCH];
  const int WP = W + 2 * PADX;
  float* s_in = reinterpret_cast<float*>(smem_raw);            // [ROWS][WP]
  const int tid = threadIdx.x;
  const int plane = blockIdx.y;
  const int y0 = blockIdx.x * BHR;
  const float* gp = g + (int64_t)plane * H * W;

  const int ya = max(y0 - R, 0), yb = min(y0 + BHR + R, H);     // valid image rows of this band
  const int top_inv = ya - (y0 - R);                            // band rows [0, top_inv) and [ROWS - bot_inv, ROWS) are padding
  const int bot_inv = (y0 + BHR + R) - yb;
  if (tid == 0) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) mbar_init(&s_bar[c], 1);
    mbar_fence_init();
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ra = max(c * CH, top_inv), rb = min(min((c + 1) * CH, ROWS), ROWS - bot_inv);
      mbar_expect_tx(&s_bar[c], (uint32_t)(max(rb - ra, 0) * W * 4));
      for (int r = ra; r < rb; ++r)
        tma_bulk_g2s(s_in + r * WP + PADX, gp + (int64_t)(y0 - R + r) * W, (uint32_t)(W * 4), &s_bar[c]);
    }
  }
  // zero margins of every row, and the rows outside the image (disjoint from every TMA destination)
  {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int MV = PADX / 4;                              // float4s per margin
    for (int e = tid; e < ROWS * 2 * MV; e += blockDim.x) {
      const int r = e / (2 * MV), q = e % (2 * MV);
      *reinterpret_cast<float4*>(s_in + r * WP + (q < MV ? 4 * q : W + PADX + 4 * (q - MV))) = z;
    }
    const int wv = W >> 2;
    for (int e = tid; e < (top_inv + bot_inv) * wv; e += blockDim.x) {
      const int k = e / wv, x4 = e - k * wv;
      const int r = k < top_inv ? k : ROWS - bot_inv + (k - top_inv);
      *reinterpret_cast<float4*>(s_in + r * WP + PADX + 4 * x4) = z;
    }
  }
  __syncthreads();     // barrier inits + zero fill visible to everyone
  if (4 * tid >= W) return;

  float wr[PW ? 1 : KS], wc[PW ? 1 : KS];
  if (!PW) {
    const int c = plane % C;
#pragma unroll
    for (int j = 0; j < KS; ++j) { wr[j] = __ldg(krow + c * KS + j); wc[j] = __ldg(kcol + c * KS + j); }
  }
  float acc[KS][4];
#pragma unroll
  for (int s = 0; s < KS; ++s) { acc[s][0] = 0.f; acc[s][1] = 0.f; acc[s][2] = 0.f; acc[s][3] = 0.f; }

  const float* sp = s_in + 4 * tid;                               // this thread's window in band row r (advanced per row)
  int yl = -(KS - 1);                                             // band-local output row completed by band row r
  const unsigned ylim = (unsigned)min(BHR, H - y0);               // rows of this band inside the image
  float* op = out + (int64_t)plane * H * W + (int64_t)(y0 + yl) * W + 4 * tid;
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    mbar_wait(&s_bar[ch], 0);
#pragma unroll
    for (int rr = 0; rr < CH; ++rr) {
      if (ch * CH + rr < ROWS) {
        const float4* row4 = reinterpret_cast<const float4*>(sp);
        float v[4 * NV];
#pragma unroll
        for (int t = 0; t < NV; ++t) {
          const float4 q = row4[t];
          v[4 * t] = q.x; v[4 * t + 1] = q.y; v[4 * t + 2] = q.z; v[4 * t + 3] = q.w;
        }
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
        for (int j = 0; j < KS; ++j) {
          const float w = PW ? wp.kr[j] : wr[j];
          t0 = fmaf(w, v[OFF + j], t0); t1 = fmaf(w, v[OFF + j + 1], t1);
          t2 = fmaf(w, v[OFF + j + 2], t2); t3 = fmaf(w, v[OFF + j + 3], t3);
        }
#pragma unroll
        for (int i = 0; i < KS; ++i) {
          const int s = ((rr - i) % KS + KS) % KS;            // static: CH is a multiple of KS
          const float w = PW ? wp.kc[i] : wc[i];
          acc[s][0] = fmaf(w, t0, acc[s][0]); acc[s][1] = fmaf(w, t1, acc[s][1]);
          acc[s][2] = fmaf(w, t2, acc[s][2]); acc[s][3] = fmaf(w, t3, acc[s][3]);
        }
        const int sc = (rr + 1) % KS;                         // slot of output row r - (KS-1): complete now
        if ((unsigned)yl < ylim)                              // false for the KS-1 warm-up rows (yl < 0) and past the image
          *reinterpret_cast<float4*>(op) = make_float4(acc[sc][0], acc[sc][1], acc[sc][2], acc[sc][3]);
        acc[sc][0] = 0.f; acc[sc][1] = 0.f; acc[sc][2] = 0.f; acc[sc][3] = 0.f;
        sp += WP; op += W; ++yl;
      }
    }
  }
}

// packed fp32x2 FMA (sm_100: SASS FFMA2, two IEEE fmas per issued instruction; same lanes per clock as FFMA — measured,
// tools/microbench/ffma_rate.cu — but half the issue slots and half the code bytes)
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float a, float b) { f32x2_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2_t p, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); }
__device__ __forceinline__ f32x2_t ffma2(f32x2_t w, f32x2_t a, f32x2_t c) {
  f32x2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(w), "l"(a), "l"(c)); return d;
}

// ---- register-sliding variant fed straight from global memory (no shared memory, no barriers) ------------------------------
// Same per-thread walk as dwconv_sep_rs_kernel, but the 4-output window of each input row comes from NV predicated 128-bit
// loads (L1-resident: neighbouring threads overlap in KS-1 of their KS+3 columns; out-of-image rows / columns are the
// predicate, i.e. the zero padding) and the next row's loads are issued before the current row's 8*KS FMAs. ncu on the
// staged kernel (profiles/ncu_tim_dim_r1b.md) showed 2.2 warps per scheduler (44 KB of staging per 2-warp CTA) and 77 % SM
// active time (1.8 waves); this form is limited by registers only and its grid is a flat list of (plane, band, column
// group) items, 128 per CTA, so that at B = 64 all of it is resident in one wave.
template <int KS, int BHR, bool PW, bool F2>
__global__ void __launch_bounds__(128, 4) dwconv_sep_rg_kernel(const float* __restrict__ g, const float* __restrict__ kcol,
                                                            const float* __restrict__ krow,
                                                            const __grid_constant__ SepWeights<KS> wp, float* __restrict__ out,
                                                            int C, int H, int W, int nbands, int64_t items, int prefetch) {
  using G = RsGeom<KS>;
  constexpr int R = G::R, PADX = G::PADX, OFF = G::OFF, NV = G::NV;
  constexpr int ROWS = BHR + KS - 1;
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= items) return;
  const int Q = W >> 2;
  const int q = (int)(item % Q);
  const int64_t pb = item / Q;
  const int band = (int)(pb % nbands);
  const int plane = (int)(pb / nbands);
  const int y0 = band * BHR;

  float wr[PW ? 1 : KS], wc[PW ? 1 : KS];
  if (!PW) {
    const int c = plane % C;
#pragma unroll
    for (int j = 0; j < KS; ++j) { wr[j] = __ldg(krow + c * KS + j); wc[j] = __ldg(kcol + c * KS + j); }
  }
  bool cv[NV];                                                     // float4 k of the window lies inside the row
#pragma unroll
  for (int k = 0; k < NV; ++k) { const int col = 4 * q - PADX + 4 * k; cv[k] = col >= 0 && col < W; }

  float acc[KS][4];
#pragma unroll
  for (int s = 0; s < KS; ++s) { acc[s][0] = 0.f; acc[s][1] = 0.f; acc[s][2] = 0.f; acc[s][3] = 0.f; }

  int yy = y0 - R;                                                 // image row of band row r
  const float4* ip = reinterpret_cast<const float4*>(g + (int64_t)plane * H * W + (int64_t)yy * W + 4 * q - PADX);
  const int pitch4 = W >> 2;
  int yl = -(KS - 1);
  const unsigned ylim = (unsigned)min(BHR, H - y0);
  float* op = out + (int64_t)plane * H * W + (int64_t)(y0 + yl) * W + 4 * q;

  float4 nxt[NV];
  {
    const bool rv = (unsigned)yy < (unsigned)H;
#pragma unroll
    for (int k = 0; k < NV; ++k) nxt[k] = (rv && cv[k]) ? __ldg(ip + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // The walk is serial per thread and only one row of loads is in flight ahead of the FMAs, so a cold input would expose one
  // DRAM latency per row (measured: 35 us at B = 64 = 46 rows x ~0.7 us). Request the whole band into L2 up front instead
  // (one line-granular prefetch per row and thread: the band's threads cover each row exactly once); the demand loads then
  // find L2 hits while DRAM streams at full rate behind them.
  if (prefetch) {
    const float* pp = g + (int64_t)plane * H * W + (int64_t)(y0 - R + 1) * W + 4 * q;
#pragma unroll 1
    for (int r = 1; r < ROWS; ++r, pp += W)
      if ((unsigned)(y0 - R + r) < (unsigned)H) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
  }
  constexpr int NG = (ROWS + KS - 1) / KS;
#pragma unroll 1
  for (int gi = 0; gi < NG; ++gi) {
#pragma unroll
    for (int rr = 0; rr < KS; ++rr) {
      if (gi * KS + rr < ROWS) {
        float v[4 * NV];
#pragma unroll
        for (int t = 0; t < NV; ++t) { v[4 * t] = nxt[t].x; v[4 * t + 1] = nxt[t].y; v[4 * t + 2] = nxt[t].z; v[4 * t + 3] = nxt[t].w; }
        ++yy; ip += pitch4;
        if (prefetch > 1 && (unsigned)(yy + 2) < (unsigned)H)      // band row r + 3 -> L1 (its L2 copy was requested up front)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(ip + 2 * pitch4 + (PADX >> 2)));
        {                                                          // prefetch band row r + 1 (predicate false past the band)
          const bool rv = (unsigned)yy < (unsigned)H && gi * KS + rr + 1 < ROWS;
#pragma unroll
          for (int k = 0; k < NV; ++k) nxt[k] = (rv && cv[k]) ? __ldg(ip + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
        if (F2) {
          // pairs along x: (t0,t1) += w_j * (v[j], v[j+1]), (t2,t3) += w_j * (v[j+2], v[j+3]); per lane the same fma chain
          f32x2_t p0 = pack2(0.f, 0.f), p1 = pack2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < KS; ++j) {
            const float w = PW ? wp.kr[j] : wr[j];
            const f32x2_t ww = pack2(w, w);
            p0 = ffma2(ww, pack2(v[OFF + j], v[OFF + j + 1]), p0);
            p1 = ffma2(ww, pack2(v[OFF + j + 2], v[OFF + j + 3]), p1);
          }
          const f32x2_t tp0 = p0, tp1 = p1;
#pragma unroll
          for (int i = 0; i < KS; ++i) {
            const int s = ((rr - i) % KS + KS) % KS;
            const float w = PW ? wp.kc[i] : wc[i];
            const f32x2_t ww = pack2(w, w);
            f32x2_t a0 = pack2(acc[s][0], acc[s][1]), a1 = pack2(acc[s][2], acc[s][3]);
            a0 = ffma2(ww, tp0, a0); a1 = ffma2(ww, tp1, a1);
            unpack2(a0, acc[s][0], acc[s][1]); unpack2(a1, acc[s][2], acc[s][3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < KS; ++j) {
            const float w = PW ? wp.kr[j] : wr[j];
            t0 = fmaf(w, v[OFF + j], t0); t1 = fmaf(w, v[OFF + j + 1], t1);
            t2 = fmaf(w, v[OFF + j + 2], t2); t3 = fmaf(w, v[OFF + j + 3], t3);
          }
#pragma unroll
          for (int i = 0; i < KS; ++i) {
            const int s = ((rr - i) % KS + KS) % KS;
            const float w = PW ? wp.kc[i] : wc[i];
            acc[s][0] = fmaf(w, t0, acc[s][0]); acc[s][1] = fmaf(w, t1, acc[s][1]);
            acc[s][2] = fmaf(w, t2, acc[s][2]); acc[s][3] = fmaf(w, t3, acc[s][3]);
          }
        }
        const int sc = (rr + 1) % KS;
        if ((unsigned)yl < ylim)
          *reinterpret_cast<float4*>(op) = make_float4(acc[sc][0], acc[sc][1], acc[sc][2], acc[sc][3]);
        acc[sc][0] = 0.f; acc[sc][1] = 0.f; acc[sc][2] = 0.f; acc[sc][3] = 0.f;
        op += W; ++yl;
      }
    }
  }
}

template <int KS, int BHR, bool PW, bool F2>
int launch_rg(const float* g, const float* kcol, const float* krow, const SepWeights<KS>& wp, float* out, int B, int C, int H,
              int W, cudaStream_t s) {
  const int nbands = (H + BHR - 1) / BHR;
  const int64_t items = (int64_t)B * C * nbands * (W / 4);
  const int64_t blocks = (items + 127) / 128;
  TA_REQUIRE(blocks <= 0x7fffffff, "ta_dwconv2d_sep: too many work items");
  dwconv_sep_rg_kernel<KS, BHR, PW, F2><<<(unsigned)blocks, 128, 0, s>>>(g, kcol, krow, wp, out, C, H, W, nbands, items,
                                                                       tune_get("tim.prefetch", 2));
  count_launch();
  return check_launch("ta_dwconv2d_sep[rg]");
}

// ---- third form of the register-sliding walk: interior / edge split, fully unrolled band ----------------------------------------
// ncu on dwconv_sep_rg_kernel<15,32> (profiles/ncu_tim_dim_r2.md): 17.3 M issued instructions of which only 6.5 M are FFMA2 —
// per input row a thread spent 60 FFMA2 + 38 MOV (building the odd-aligned operand pairs v[j], v[j+1] of the row pass) + 10 CS2R
// (zero-filling the destination of its five predicated loads) + ~35 integer / predicate instructions, and ran all KS column taps
// on the 2 x (KS - 1) halo rows although a halo row feeds only part of the band. Here:
//  * row pass with the DATA broadcast and the WEIGHTS paired: (out[x], out[x+1]) += (w[m], w[m-1]) * v[x+m] — SASS
//    `FFMA2 R, R.F32, UR.F32x2, R`: the pair operand is a uniform-register pair of kernel parameters, no per-row register moves;
//    per output the products still arrive in tap order 0..KS-1 from +0 (the end taps of a pair are scalar FFMAs), so the result
//    is the same fma chain as before, bit for bit;
//  * the band's BHR + KS - 1 rows are unrolled completely, so which column taps a row feeds (i <= r at the top, i >= r - BHR + 1 at
//    the bottom) is decided at compile time: 2 * BHR * KS column FFMA2s per thread instead of 2 * (BHR + KS - 1) * KS, and the
//    first tap of every output row takes +0 as its addend instead of a zeroed accumulator;
//  * threads whose 4-output window (KS + 3 columns, as NV 128-bit loads) lies inside the row — all but 2 + 2 per row at ks = 15 —
//    run in their own CTAs with unconditional loads; the edge windows get CTAs of their own with the predicated form. Rows outside
//    the image are skipped (their products are exact zeros and an accumulator is never -0).
template <int KS> struct SepWeights2 { float kr[KS]; float kc[KS]; float wp[KS + 1][2]; };   // wp[m] = (w[m], w[m-1]), w[-1] = w[KS] = 0

__device__ __forceinline__ f32x2_t ffma2_bc(float2 wpair, float v, f32x2_t c) {   // (c.lo, c.hi) + (wpair.x, wpair.y) * v
  return ffma2(pack2(v, v), pack2(wpair.x, wpair.y), c);
}

// Zero rows for the constant-width walk: a window quarter that lies outside the image row reads HERE instead (same row stride, so
// the unrolled walk keeps its [base + immediate] addressing and every load is unconditional). 64 rows x 224 floats of zeros.
__device__ __align__(16) float g_rg2_zero_rows[64 * 224 + 8];

// 128-bit read-only load under a predicate that leaves the destination registers untouched when false: a window quarter that lies
// outside the row keeps the zeros it was initialised with for the whole walk (no per-row zero fill, no select)
__device__ __forceinline__ void ldg4_if(float4& d, const float4* p, int pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
               : "+f"(d.x), "+f"(d.y), "+f"(d.z), "+f"(d.w) : "l"(p), "r"(pred));
}

// WC > 0: the image width is a compile-time constant (224, the hot shape): every load, store and prefetch of the walk is
// [base register + immediate] — no per-row pointer arithmetic and no constant-bank reads of W behind a scoreboard.
template <int KS, int BHR, bool EDGE, int WC, bool DEEP>
__device__ __forceinline__ void rg2_walk(const float* __restrict__ g, float* __restrict__ out, const SepWeights2<KS>& wp,
                                         int plane, int band, int nbands, int q, int H, int W_rt, int prefetch) {
  const int W = WC ? WC : W_rt;
  using G = RsGeom<KS>;
  constexpr int R = G::R, PADX = G::PADX, OFF = G::OFF, NV = G::NV;
  constexpr int ROWS = BHR + KS - 1;
  const int y0 = band * BHR;                              // H % BHR == 0 (host-checked): every band is full, only the first R rows of
  const bool top_ok = band > 0, bot_ok = band < nbands - 1;   // band 0 and the last R rows of the last band lie outside the image
  int cv[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) { const int col = 4 * q - PADX + 4 * k; cv[k] = (!EDGE || (col >= 0 && col < W)) ? 1 : 0; }
  const float4* ip0 = reinterpret_cast<const float4*>(g + (int64_t)plane * H * W + (int64_t)(y0 - R) * W + 4 * q - PADX);
  float* op0 = out + (int64_t)plane * H * W + (int64_t)y0 * W + 4 * q;
  const int pitch4 = W >> 2;
  f32x2_t acc0[KS], acc1[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) { acc0[s] = pack2(0.f, 0.f); acc1[s] = pack2(0.f, 0.f); }
  float4 buf[2][NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) { buf[0][k] = make_float4(0.f, 0.f, 0.f, 0.f); buf[1][k] = make_float4(0.f, 0.f, 0.f, 0.f); }
  constexpr bool ZB = EDGE && WC == 224 && ROWS <= 64;      // out-of-row window quarters read the zero rows: no predicates at all
  const float4* bk[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) bk[k] = (!ZB || cv[k]) ? ip0 + k : reinterpret_cast<const float4*>(g_rg2_zero_rows);
#define TA_RG2_LOAD(SLOT_, ROW_)                                                               \
  do {                                                                                         \
    _Pragma("unroll") for (int k = 0; k < NV; ++k) {                                           \
      if (ZB) buf[SLOT_][k] = __ldg(bk[k] + (ROW_) * pitch4);                                  \
      else if (EDGE) ldg4_if(buf[SLOT_][k], ip0 + (ROW_) * pitch4 + k, cv[k]);                 \
      else buf[SLOT_][k] = __ldg(ip0 + (ROW_) * pitch4 + k);                                   \
    }                                                                                          \
  } while (0)
#define TA_RG2_LOAD_ROW(ROW_)                                                                  \
  do {                                                                                         \
    if ((ROW_) < ROWS) {                                                                       \
      if ((ROW_) < R) { if (top_ok) TA_RG2_LOAD((ROW_) & 1, ROW_); }                           \
      else if ((ROW_) >= ROWS - R) { if (bot_ok) TA_RG2_LOAD((ROW_) & 1, ROW_); }              \
      else TA_RG2_LOAD((ROW_) & 1, ROW_);                                                      \
    }                                                                                          \
  } while (0)
  // DEEP: the loads of row r + 2 go out as soon as the row pass of row r has consumed its buffer (before the column pass), i.e.
  // ~1.4 row-times ahead of their use instead of 1.0
  TA_RG2_LOAD_ROW(0);
  if (DEEP) TA_RG2_LOAD_ROW(1);
  // The walk is serial per thread with one row of loads in flight ahead of the FMAs, i.e. a cold row costs one DRAM latency.
  // prefetch 3 (default): every row step asks L2 for this thread's own 16 bytes of the row PD rows further down (the band's
  // threads cover each row once) — a rolling request stream PD row-times ahead of the demand loads instead of one burst;
  // 1 / 2: the whole band up front (the earlier kernels' scheme; measured slower here), 2 also the row three ahead into L1.
  constexpr int PD = 6;
  if (prefetch == 1 || prefetch == 2) {
    const float* pp = g + (int64_t)plane * H * W + (int64_t)(y0 - R + 1) * W + 4 * q;
#pragma unroll 1
    for (int r = 1; r < ROWS; ++r, pp += W)
      if ((unsigned)(y0 - R + r) < (unsigned)H) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
  } else if (prefetch == 3) {
#pragma unroll
    for (int r = 2; r < PD; ++r)
      if ((r >= R || top_ok) && (r < ROWS - R || bot_ok)) asm volatile("prefetch.global.L2 [%0];" ::"l"(ip0 + r * pitch4 + (PADX >> 2)));
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (prefetch == 2 && r + 3 < ROWS && (r + 3 >= R || top_ok) && (r + 3 < ROWS - R || bot_ok))
      asm volatile("prefetch.global.L1 [%0];" ::"l"(ip0 + (r + 3) * pitch4 + (PADX >> 2)));
    if (prefetch == 3 && r + PD < ROWS && (r + PD >= R || top_ok) && (r + PD < ROWS - R || bot_ok))
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ip0 + (r + PD) * pitch4 + (PADX >> 2)));
    if (!DEEP) TA_RG2_LOAD_ROW(r + 1);
    const bool rv = r < R ? top_ok : (r >= ROWS - R ? bot_ok : true);
    if (rv) {
      float v[4 * NV];
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        v[4 * t] = buf[r & 1][t].x; v[4 * t + 1] = buf[r & 1][t].y; v[4 * t + 2] = buf[r & 1][t].z; v[4 * t + 3] = buf[r & 1][t].w;
      }
      // row pass: pair 0 = outputs (0, 1), pair 1 = outputs (2, 3); tap index m, data v[OFF + m (+ 2)]
      float lo0 = fmaf(wp.kr[0], v[OFF], 0.f), lo1 = fmaf(wp.kr[0], v[OFF + 2], 0.f);
      f32x2_t p0 = pack2(lo0, 0.f), p1 = pack2(lo1, 0.f);
#pragma unroll
      for (int m = 1; m < KS; ++m) {
        const float2 w2 = make_float2(wp.wp[m][0], wp.wp[m][1]);
        p0 = ffma2_bc(w2, v[OFF + m], p0);
        p1 = ffma2_bc(w2, v[OFF + 2 + m], p1);
      }
      float a0, a1, b0, b1;
      unpack2(p0, a0, a1); unpack2(p1, b0, b1);
      a1 = fmaf(wp.kr[KS - 1], v[OFF + KS], a1);
      b1 = fmaf(wp.kr[KS - 1], v[OFF + 2 + KS], b1);
      const f32x2_t t0 = pack2(a0, a1), t1 = pack2(b0, b1);
      if (DEEP) TA_RG2_LOAD_ROW(r + 2);
      // column pass: this row is tap i of output row y = r - i
#pragma unroll
      for (int i = 0; i < KS; ++i) {
        const int y = r - i;
        if (y >= 0 && y < BHR) {
          const int s = y % KS;
          const f32x2_t ww = pack2(wp.kc[i], wp.kc[i]);
          if (i == 0) { acc0[s] = ffma2(ww, t0, pack2(0.f, 0.f)); acc1[s] = ffma2(ww, t1, pack2(0.f, 0.f)); }
          else { acc0[s] = ffma2(ww, t0, acc0[s]); acc1[s] = ffma2(ww, t1, acc1[s]); }
        }
      }
    } else {
      if (DEEP) TA_RG2_LOAD_ROW(r + 2);
      if (r < BHR) {                                       // the first tap of output row r never comes: start it at +0
        acc0[r % KS] = pack2(0.f, 0.f); acc1[r % KS] = pack2(0.f, 0.f);
      }
    }
    if (r >= KS - 1) {
      const int s = (r - (KS - 1)) % KS;
      float o0, o1, o2, o3;
      unpack2(acc0[s], o0, o1); unpack2(acc1[s], o2, o3);
      *reinterpret_cast<float4*>(op0 + (r - (KS - 1)) * W) = make_float4(o0, o1, o2, o3);
    }
  }
#undef TA_RG2_LOAD_ROW
#undef TA_RG2_LOAD
}

// ---- fourth form: the same walk fed from a warp-private shared-memory ring (cp.async), W = 224 ---------------------------------
// ncu on the walk above: 59 % of the stall samples sit on the first FFMA of a row, waiting for that row's global loads — with 128
// registers a thread can hold only one row of loads in flight ahead of its FMAs, and 4 warps per scheduler do not cover an L2
// round trip per row. Here the prefetch depth is decoupled from the register file: a WARP owns 28 adjacent windows (112
// output columns + 16 halo columns = exactly 32 x 16 bytes per input row); every lane copies one 16-byte piece of each row with
// cp.async into a ring of NR rows in the warp's own shared memory, NR - 1 rows ahead of the row being consumed; a row is then
// read back as five conflict-free LDS.128 (lane l: pieces l .. l + 4). No CTA barrier (warp-level wait_group + __syncwarp), no
// redundant global loads (each element is fetched once per warp instead of five times), 20 registers less per thread. Lanes
// 28-31 only copy. Out-of-row pieces (two per edge warp) read the zero rows like above; same FMA chains → same bits.
// MEASURED (B = 64): 26.6 us against 24.6 us for the walk above — the memory stalls are gone (long_scoreboard 4.7 → 0.6 per issue)
// and `no_instruction` takes their place (2.3 per issue): 20 warps per SM each stream 60 KB of straight-line code through a
// 32 KB L1.5 / 6 KB L0 instruction cache. Rolling the walk into 15-row groups (run-time ring slots, all taps on every row) made
// ptxas rotate the accumulator file through moves and spill: 110 us; removed. Kept behind tim.band = 5 as the starting point
// for a version with a smaller code footprint.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kRg3Warps = 2;          // warps per CTA = windows of 28 x 4 columns per 224-column row
constexpr int kRg3NR = 8;             // ring depth (rows)

template <int KS, int BHR>
__global__ void __launch_bounds__(32 * kRg3Warps, 10) dwconv_sep_rg3_kernel(const float* __restrict__ g, const __grid_constant__ SepWeights2<KS> wp,
                                                                        float* __restrict__ out, int H, int nbands) {
  using G = RsGeom<KS>;
  constexpr int R = G::R, PADX = G::PADX, OFF = G::OFF, NV = G::NV;
  constexpr int W = 224, ROWS = BHR + KS - 1, NR = kRg3NR, LW = 28;          // LW: windows (threads that compute) per warp
  static_assert(KS == 15 && NV == 5 && PADX == 8, "the ring layout is written for ks = 15");
  __shared__ __align__(16) float ring[kRg3Warps][NR * 128 + 16];            // + 16: lanes 28-31 read 4 pieces past a row (discarded)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int band = blockIdx.x % nbands, plane = blockIdx.x / nbands;
  const int y0 = band * BHR;
  const bool top_ok = band > 0, bot_ok = band < nbands - 1;
  // this lane's 16-byte piece of every input row: columns [c0, c0 + 4), c0 = 112 * warp - 8 + 4 * lane
  const int c0 = LW * 4 * warp - PADX + 4 * lane;
  const float* src0 = (c0 >= 0 && c0 < W) ? g + (int64_t)plane * H * W + (int64_t)(y0 - R) * W + c0 : g_rg2_zero_rows;
  float* my = ring[warp];
  float* dst0 = my + 4 * lane;
  float* op0 = out + (int64_t)plane * H * W + (int64_t)y0 * W + (LW * 4 * warp + 4 * lane);
  const bool writer = lane < LW;

  f32x2_t acc0[KS], acc1[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) { acc0[s] = pack2(0.f, 0.f); acc1[s] = pack2(0.f, 0.f); }

#define TA_RG3_COPY(ROW_)                                                                                     \
  do {                                                                                                        \
    if ((ROW_) < ROWS) {                                                                                      \
      const bool ok_ = (ROW_) < R ? top_ok : ((ROW_) >= ROWS - R ? bot_ok : true);                            \
      if (ok_) cp_async16(dst0 + ((ROW_) % NR) * 128, src0 + (ROW_) * W);                                     \
    }                                                                                                         \
    cp_async_commit();                                                                                        \
  } while (0)
#pragma unroll
  for (int r = 0; r < NR - 1; ++r) TA_RG3_COPY(r);

#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    cp_async_wait<NR - 2>();                              // row r has landed (this lane's piece)
    __syncwarp();                                         // ... and every other lane's; all lanes are done reading row r - 1
    TA_RG3_COPY(r + NR - 1);                              // refill the slot row r - 1 occupied
    const bool rv = r < R ? top_ok : (r >= ROWS - R ? bot_ok : true);
    if (rv) {
      float v[4 * NV];
      const float4* rp = reinterpret_cast<const float4*>(my + (r % NR) * 128) + lane;
#pragma unroll
      for (int t = 0; t < NV; ++t) { const float4 x = rp[t]; v[4 * t] = x.x; v[4 * t + 1] = x.y; v[4 * t + 2] = x.z; v[4 * t + 3] = x.w; }
      float lo0 = fmaf(wp.kr[0], v[OFF], 0.f), lo1 = fmaf(wp.kr[0], v[OFF + 2], 0.f);
      f32x2_t p0 = pack2(lo0, 0.f), p1 = pack2(lo1, 0.f);
#pragma unroll
      for (int m = 1; m < KS; ++m) {
        const float2 w2 = make_float2(wp.wp[m][0], wp.wp[m][1]);
        p0 = ffma2_bc(w2, v[OFF + m], p0);
        p1 = ffma2_bc(w2, v[OFF + 2 + m], p1);
      }
      float a0, a1, b0, b1;
      unpack2(p0, a0, a1); unpack2(p1, b0, b1);
      a1 = fmaf(wp.kr[KS - 1], v[OFF + KS], a1);
      b1 = fmaf(wp.kr[KS - 1], v[OFF + 2 + KS], b1);
      const f32x2_t t0 = pack2(a0, a1), t1 = pack2(b0, b1);
#pragma unroll
      for (int i = 0; i < KS; ++i) {
        const int y = r - i;
        if (y >= 0 && y < BHR) {
          const int s = y % KS;
          const f32x2_t ww = pack2(wp.kc[i], wp.kc[i]);
          if (i == 0) { acc0[s] = ffma2(ww, t0, pack2(0.f, 0.f)); acc1[s] = ffma2(ww, t1, pack2(0.f, 0.f)); }
          else { acc0[s] = ffma2(ww, t0, acc0[s]); acc1[s] = ffma2(ww, t1, acc1[s]); }
        }
      }
    } else if (r < BHR) {
      acc0[r % KS] = pack2(0.f, 0.f); acc1[r % KS] = pack2(0.f, 0.f);
    }
    if (r >= KS - 1) {
      const int s = (r - (KS - 1)) % KS;
      float o0, o1, o2, o3;
      unpack2(acc0[s], o0, o1); unpack2(acc1[s], o2, o3);
      if (writer) *reinterpret_cast<float4*>(op0 + (r - (KS - 1)) * W) = make_float4(o0, o1, o2, o3);
    }
  }
  cp_async_wait<0>();
#undef TA_RG3_COPY
}

template <int KS, int BHR>
int launch_rg3(const float* g, const float* kcol_host, const float* krow_host, float* out, int B, int C, int H, cudaStream_t s) {
  SepWeights2<KS> w;
  for (int j = 0; j < KS; ++j) { w.kr[j] = krow_host[j]; w.kc[j] = kcol_host[j]; }
  for (int m = 0; m <= KS; ++m) { w.wp[m][0] = m < KS ? krow_host[m] : 0.0f; w.wp[m][1] = m >= 1 ? krow_host[m - 1] : 0.0f; }
  const int nbands = H / BHR;
  const int64_t blocks = (int64_t)B * C * nbands;
  TA_REQUIRE(blocks <= 0x7fffffff, "ta_dwconv2d_sep: too many work items");
  dwconv_sep_rg3_kernel<KS, BHR><<<(unsigned)blocks, 32 * kRg3Warps, 0, s>>>(g, w, out, H, nbands);
  count_launch();
  return check_launch("ta_dwconv2d_sep[rg3]");
}

// SPLIT = false (default): one code path, every thread with the predicated loads. SPLIT = true: the interior windows in CTAs of
// their own with unconditional loads, the edge windows in trailing CTAs — measured slower: the few edge warps stream 70 KB of
// straight-line code that no other warp on their SM has brought into the instruction cache (ncu: 83 % of their stall samples
// are no_instruction) and run 4x longer than the interior warps.
template <int KS, int BHR, bool SPLIT, int WC, bool DEEP>
__global__ void __launch_bounds__(128, 4) dwconv_sep_rg2_kernel(const float* __restrict__ g, const __grid_constant__ SepWeights2<KS> wp,
                                                             float* __restrict__ out, int H, int W, int nbands, int n_int_blocks,
                                                             int64_t n_int, int64_t n_edge, int prefetch) {
  using G = RsGeom<KS>;
  constexpr int NEL = G::PADX / 4, NER = G::NV - 1 - G::PADX / 4;     // edge windows per row, left / right
  const int Q = W >> 2;
  if (!SPLIT) {
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_int) return;
    const int q = (int)(item % Q);
    const int64_t pb = item / Q;
    rg2_walk<KS, BHR, true, WC, DEEP>(g, out, wp, (int)(pb / nbands), (int)(pb % nbands), nbands, q, H, W, prefetch);
  } else if ((int)blockIdx.x < n_int_blocks) {
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_int) return;
    const int QI = Q - NEL - NER;
    const int q = NEL + (int)(item % QI);
    const int64_t pb = item / QI;
    rg2_walk<KS, BHR, false, WC, DEEP>(g, out, wp, (int)(pb / nbands), (int)(pb % nbands), nbands, q, H, W, prefetch);
  } else {
    const int64_t item = (int64_t)(blockIdx.x - n_int_blocks) * blockDim.x + threadIdx.x;
    if (item >= n_edge) return;
    const int e = (int)(item % (NEL + NER));
    const int64_t pb = item / (NEL + NER);
    const int q = e < NEL ? e : Q - (NEL + NER) + e;
    rg2_walk<KS, BHR, true, WC, DEEP>(g, out, wp, (int)(pb / nbands), (int)(pb % nbands), nbands, q, H, W, prefetch);
  }
}

template <int KS, int BHR>
int launch_rg2(const float* g, const float* kcol_host, const float* krow_host, float* out, int B, int C, int H, int W, cudaStream_t s) {
  using G = RsGeom<KS>;
  constexpr int NE = G::NV - 1;
  SepWeights2<KS> w;
  for (int j = 0; j < KS; ++j) { w.kr[j] = krow_host[j]; w.kc[j] = kcol_host[j]; }
  for (int m = 0; m <= KS; ++m) { w.wp[m][0] = m < KS ? krow_host[m] : 0.0f; w.wp[m][1] = m >= 1 ? krow_host[m - 1] : 0.0f; }
  const int nbands = (H + BHR - 1) / BHR, Q = W / 4;
  const int64_t pbs = (int64_t)B * C * nbands;
  const int pf = tune_get("tim.prefetch2", 3);
  if (tune_get("tim.split", 0) != 0) {
    const int64_t n_int = pbs * (Q - NE), n_edge = pbs * NE;
    const int64_t bi = (n_int + 127) / 128, be = (n_edge + 127) / 128;
    TA_REQUIRE(bi + be <= 0x7fffffff, "ta_dwconv2d_sep: too many work items");
    dwconv_sep_rg2_kernel<KS, BHR, true, 0, false><<<(unsigned)(bi + be), 128, 0, s>>>(g, w, out, H, W, nbands, (int)bi, n_int, n_edge, pf);
  } else {
    const int64_t items = pbs * Q;
    const int64_t blocks = (items + 127) / 128;
    TA_REQUIRE(blocks <= 0x7fffffff, "ta_dwconv2d_sep: too many work items");
    if (W == 224 && tune_get("tim.wconst", 1) != 0) {
      if (tune_get("tim.deep", 0) != 0)       // loads two rows ahead: same time, but 128 registers then spill (ptxas: 28 B)
        dwconv_sep_rg2_kernel<KS, BHR, false, 224, true><<<(unsigned)blocks, 128, 0, s>>>(g, w, out, H, W, nbands, 0, items, 0, pf);
      else
        dwconv_sep_rg2_kernel<KS, BHR, false, 224, false><<<(unsigned)blocks, 128, 0, s>>>(g, w, out, H, W, nbands, 0, items, 0, pf);
    } else {
      dwconv_sep_rg2_kernel<KS, BHR, false, 0, false><<<(unsigned)blocks, 128, 0, s>>>(g, w, out, H, W, nbands, 0, items, 0, pf);
    }
  }
  count_launch();
  return check_launch("ta_dwconv2d_sep[rg2]");
}

template <int KS, int BHR, bool PW>
int launch_rs(const float* g, const float* kcol, const float* krow, const SepWeights<KS>& wp, float* out, int B, int C, int H,
              int W, cudaStream_t s) {
  using G = RsGeom<KS>;
  const size_t smem = sizeof(float) * (size_t)(BHR + KS - 1) * (W + 2 * G::PADX);
  auto k = dwconv_sep_rs_kernel<KS, BHR, PW>;
  static SmemOptIn optin = {};
  const int rc = ensure_dyn_smem("ta_dwconv2d_sep", k, smem, optin);
  if (rc != TA_OK) return rc;
  const int threads = (((W + 3) / 4) + 31) & ~31;
  dim3 grid((unsigned)((H + BHR - 1) / BHR), (unsigned)(B * C));
  k<<<grid, threads, smem, s>>>(g, kcol, krow, wp, out, C, H, W);
  count_launch();
  return check_launch("ta_dwconv2d_sep[rs]");
}

template <int KS, bool PW>
int launch_rs_bh(const float* g, const float* kcol, const float* krow, const SepWeights<KS>& wp, float* out, int B, int C,
                 int H, int W, cudaStream_t s) {
  const int bh = tune_get("tim.bh", 32);
  if (tune_get("tim.band", 4) >= 3) {       // straight from global memory; tim.f2: packed fp32x2 FMAs
    if (tune_get("tim.f2", 1) != 0) {
      if (bh == 56) return launch_rg<KS, 56, PW, true>(g, kcol, krow, wp, out, B, C, H, W, s);
      return launch_rg<KS, 32, PW, true>(g, kcol, krow, wp, out, B, C, H, W, s);
    }
    if (bh == 56) return launch_rg<KS, 56, PW, false>(g, kcol, krow, wp, out, B, C, H, W, s);
    return launch_rg<KS, 32, PW, false>(g, kcol, krow, wp, out, B, C, H, W, s);
  }
  if (bh == 56) return launch_rs<KS, 56, PW>(g, kcol, krow, wp, out, B, C, H, W, s);
  return launch_rs<KS, 32, PW>(g, kcol, krow, wp, out, B, C, H, W, s);
}

inline bool rs_ok(const void* g, const void* out, int ks, int W) {
  return (W % 4 == 0) && W >= 32 && W <= 512 && aligned16(g) && aligned16(out) && (ks == 3 || ks == 5 || ks == 7 || ks == 15);
}

template <int KS>
__global__ void __launch_bounds__(kThreads) dwconv2d_kernel(const float* __restrict__ g, const float* __restrict__ k, int ks_rt,
                                                            float* __restrict__ out, int C, int H, int W) {
  extern __shared__ __align__(16) float smem[];
  const int ks = KS > 0 ? KS : ks_rt;
  const int th = TH + ks - 1;
  const int IS = odd_up(TW + ks - 1);
  float* s_in = smem;                 // [th][IS]
  float* s_k = s_in + th * IS;        // [ks*ks]

  const int plane = blockIdx.z, c = plane % C;
  const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
  const float* gp = g + (int64_t)plane * H * W;
  float* op = out + (int64_t)plane * H * W;
  const int tid = threadIdx.x;
  for (int e = tid; e < ks * ks; e += kThreads) s_k[e] = __ldg(k + (int64_t)c * ks * ks + e);
  load_tile(gp, s_in, IS, ks, H, W, y0, x0);
  __syncthreads();

  // each thread: 4 consecutive outputs of one row; 32 rows x 8 groups = 256 items
  const int y = tid / (TW / 4), xg = tid % (TW / 4);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (KS > 0) {
#pragma unroll 1
    for (int i = 0; i < KS; ++i) {
      const float* row = s_in + (y + i) * IS + 4 * xg;
      float v[(KS > 0 ? KS : 1) + 3];
#pragma unroll
      for (int t = 0; t < KS + 3; ++t) v[t] = row[t];
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const float w = s_k[i * KS + j];
        a0 = fmaf(w, v[j], a0); a1 = fmaf(w, v[j + 1], a1); a2 = fmaf(w, v[j + 2], a2); a3 = fmaf(w, v[j + 3], a3);
      }
    }
  } else {
    for (int i = 0; i < ks; ++i) {
      const float* row = s_in + (y + i) * IS + 4 * xg;
      for (int j = 0; j < ks; ++j) {
        const float w = s_k[i * ks + j];
        a0 = fmaf(w, row[j], a0); a1 = fmaf(w, row[j + 1], a1); a2 = fmaf(w, row[j + 2], a2); a3 = fmaf(w, row[j + 3], a3);
      }
    }
  }
  const int yy = y0 + y;
  if (yy < H) {
    const int xx = x0 + 4 * xg;
    float* o = op + (int64_t)yy * W + xx;
    if (xx < W) o[0] = a0;
    if (xx + 1 < W) o[1] = a1;
    if (xx + 2 < W) o[2] = a2;
    if (xx + 3 < W) o[3] = a3;
  }
}

int check_conv(const char* who, const void* g, const void* k, const void* out, int ks, int B, int C, int H, int W) {
  TA_REQUIRE(g && k && out, "%s: null pointer", who);
  TA_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "%s: empty shape", who);
  TA_REQUIRE(ks >= 1 && (ks & 1) == 1, "%s: kernel size %d must be odd ('same' padding)", who, ks);
  if (ks > kMaxKs) { set_error("%s: kernel size %d > %d not supported", who, ks, kMaxKs); return TA_EUNSUPPORTED; }
  TA_REQUIRE((int64_t)B * C <= 65535, "%s: B*C=%lld exceeds 65535 planes", who, (long long)B * C);
  return TA_OK;
}

}  // namespace

extern "C" {

int ta_dwconv2d_sep(const float* g, const float* kcol, const float* krow, int ks, float* out, int B, int C, int H, int W,
                    ta_stream_t stream) {
  int rc = check_conv("ta_dwconv2d_sep", g, kcol, out, ks, B, C, H, W);
  if (rc != TA_OK) return rc;
  TA_REQUIRE(krow, "ta_dwconv2d_sep: null krow");
  // hot case (TIM on 224-class images), tim.band: 3 = register-sliding kernel fed from global memory (default), 2 = the
  // same fed from bulk-TMA-staged shared memory, 1 = two-pass band kernel, 0 = 32x32 tiles
  const int mode = tune_get("tim.band", 3);
  if (mode >= 2 && rs_ok(g, out, ks, W)) {
    cudaStream_t bs = (cudaStream_t)stream;
    switch (ks) {
      case 3: return launch_rs_bh<3, false>(g, kcol, krow, SepWeights<3>{}, out, B, C, H, W, bs);
      case 5: return launch_rs_bh<5, false>(g, kcol, krow, SepWeights<5>{}, out, B, C, H, W, bs);
      case 7: return launch_rs_bh<7, false>(g, kcol, krow, SepWeights<7>{}, out, B, C, H, W, bs);
      default: return launch_rs_bh<15, false>(g, kcol, krow, SepWeights<15>{}, out, B, C, H, W, bs);
    }
  }
  if ((W % 4 == 0) && W >= 32 && W <= 512 && aligned16(g) && mode != 0) {
    cudaStream_t bs = (cudaStream_t)stream;
    switch (ks) {
      case 3: return launch_band<3>(g, kcol, krow, out, B, C, H, W, bs);
      case 5: return launch_band<5>(g, kcol, krow, out, B, C, H, W, bs);
      case 7: return launch_band<7>(g, kcol, krow, out, B, C, H, W, bs);
      case 15: return launch_band<15>(g, kcol, krow, out, B, C, H, W, bs);
      default: break;
    }
  }
  const int th = TH + ks - 1, IS = (TW + ks - 1) | 1;
  const size_t smem = sizeof(float) * ((size_t)th * IS + (size_t)th * TW + 2 * kMaxKs);
  dim3 grid((unsigned)((W + TW - 1) / TW), (unsigned)((H + TH - 1) / TH), (unsigned)(B * C));
  cudaStream_t s = (cudaStream_t)stream;
#define TA_SEP_CASE(K)                                                                          \
  case K:                                                                                       \
    dwconv_sep_kernel<K><<<grid, kThreads, smem, s>>>(g, kcol, krow, ks, out, C, H, W);         \
    break;
  switch (ks) {
    TA_SEP_CASE(3) TA_SEP_CASE(5) TA_SEP_CASE(7) TA_SEP_CASE(9) TA_SEP_CASE(11) TA_SEP_CASE(15)
    default:
      dwconv_sep_kernel<0><<<grid, kThreads, smem, s>>>(g, kcol, krow, ks, out, C, H, W);
  }
#undef TA_SEP_CASE
  count_launch();
  return check_launch("ta_dwconv2d_sep");
}

// Host-weight form: kcol_host / krow_host are HOST arrays [C, ks] read during the call (like ta_lin_sample_fwd's table).
// When every channel carries the same factors (all of tim.py's kernels) and the shape is the hot one, the weights travel
// as kernel parameters and feed the FMAs from the constant bank; otherwise returns TA_EUNSUPPORTED and the caller uses
// ta_dwconv2d_sep with device arrays.
int ta_dwconv2d_sep_hw(const float* g, const float* kcol_host, const float* krow_host, int ks, float* out, int B, int C,
                       int H, int W, ta_stream_t stream) {
  int rc = check_conv("ta_dwconv2d_sep_hw", g, kcol_host, out, ks, B, C, H, W);
  if (rc != TA_OK) return rc;
  TA_REQUIRE(krow_host, "ta_dwconv2d_sep_hw: null krow_host");
  bool same = true;
  for (int c = 1; c < C && same; ++c)
    for (int j = 0; j < ks; ++j)
      if (memcmp(&kcol_host[c * ks + j], &kcol_host[j], 4) != 0 || memcmp(&krow_host[c * ks + j], &krow_host[j], 4) != 0) { same = false; break; }
  if (!same || !rs_ok(g, out, ks, W)) {
    set_error("ta_dwconv2d_sep_hw: needs channel-shared factors, ks in {3,5,7,15}, W %% 4 == 0, 32 <= W <= 512, 16-B aligned tensors");
    return TA_EUNSUPPORTED;
  }
  cudaStream_t bs = (cudaStream_t)stream;
  const int band_mode = tune_get("tim.band", 4);
  if (band_mode == 5 && ks == 15 && W == 224 && H % 32 == 0 && H >= 64)                   // zero rows cover 46 rows of 224
    return launch_rg3<15, 32>(g, kcol_host, krow_host, out, B, C, H, bs);
#define TA_HW_CASE(K)                                                                     \
  case K: {                                                                               \
    if (band_mode >= 4 && W / 4 > RsGeom<K>::NV - 1 && H % 32 == 0 && H >= 64)            \
      return launch_rg2<K, 32>(g, kcol_host, krow_host, out, B, C, H, W, bs);             \
    SepWeights<K> w;                                                                      \
    for (int j = 0; j < K; ++j) { w.kr[j] = krow_host[j]; w.kc[j] = kcol_host[j]; }      \
    return launch_rs_bh<K, true>(g, nullptr, nullptr, w, out, B, C, H, W, bs);            \
  }
  switch (ks) {
    TA_HW_CASE(3) TA_HW_CASE(5) TA_HW_CASE(7) TA_HW_CASE(15)
    default: break;
  }
#undef TA_HW_CASE
  return TA_EUNSUPPORTED;
}

int ta_dwconv2d(const float* g, const float* k, int ks, float* out, int B, int C, int H, int W, ta_stream_t stream) {
  int rc = check_conv("ta_dwconv2d", g, k, out, ks, B, C, H, W);
  if (rc != TA_OK) return rc;
  const int th = TH + ks - 1, IS = (TW + ks - 1) | 1;
  const size_t smem = sizeof(float) * ((size_t)th * IS + (size_t)ks * ks);
  dim3 grid((unsigned)((W + TW - 1) / TW), (unsigned)((H + TH - 1) / TH), (unsigned)(B * C));
  cudaStream_t s = (cudaStream_t)stream;
#define TA_2D_CASE(K)                                                                  \
  case K:                                                                              \
    dwconv2d_kernel<K><<<grid, kThreads, smem, s>>>(g, k, ks, out, C, H, W);           \
    break;
  switch (ks) {
    TA_2D_CASE(3) TA_2D_CASE(5) TA_2D_CASE(7) TA_2D_CASE(15)
    default:
      dwconv2d_kernel<0><<<grid, kThreads, smem, s>>>(g, k, ks, out, C, H, W);
  }
#undef TA_2D_CASE
  count_launch();
  return check_launch("ta_dwconv2d");
}

}  // extern "C"
