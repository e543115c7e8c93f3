// spectrum.cu — SSM / FGSRA spectrum transform (input_transformation/ssm.py:41-55: x_idct = idct_2d(dct_2d(x + gauss) * mask),
// dct / idct at ssm.py:101-200) as FOUR tensor-core GEMMs on tcgen05 (the one dense contraction on the path).
//
// The reference evaluates the 224-point DCT-II / its inverse along rows and columns through FFTs (≈ 40 ATen launches and
// ≈ 4 GB of traffic per transform at B = 64). Written as matrices, with D[k][n] = 2 cos(pi (2n+1) k / 2N) and E = D^-1:
//     T(X) = E ((D X D^T) . M) E^T                                  per (sample, channel) plane X [N x N]
// Every factor is the same primitive  P(A; W) = (A W^T)^T = W A^T  applied to all planes at once:
//     R1 = P(X + gauss; D) = D X^T,   Y = P(R1; D) . M = (D X D^T) . M,   R3 = P(Y; E),   T = P(R3; E).
// P is one kernel: C[r][n] = sum_k A[r][k] W[n][k] over the stacked rows r = (plane, i) of all planes (a [planes*N, N] x [N, N]
// GEMM with both operands K-major), epilogue writes C transposed inside its plane (out[plane][n][i]) times an optional mask.
//
// tcgen05 mapping (one CTA = 128 stacked rows x all N columns; K in blocks of 32):
//   * operands are staged by the CTA's threads from global memory into shared memory in the canonical K-major NO-SWIZZLE UMMA
//     layout (8-row x 16-byte core matrices; LBO = 128 B between the K-adjacent cores, SBO = 1 KB between 8-row groups), and
//     SPLIT on the way: v = hi + lo with hi = v truncated to tf32's 10-bit mantissa (exactly what the tensor core would keep)
//     and lo = v - hi (exact). The (x + gauss) add of stage 1 happens in the same pass;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = N, K = 8) three times per K step:
//     hi*hi + lo*hi + hi*lo — "3xTF32": the dropped lo*lo term and tf32's truncation of lo are ~2^-21 relative, so the fp32
//     accumulator in TMEM carries fp32-level products (measured against an fp64 restatement in the tests);
//   * tcgen05.commit -> mbarrier tells the CTA when the staged block may be overwritten and when the accumulator is complete;
//   * epilogue: tcgen05.ld (32 lanes x 32-bit x 16 columns per warp and step) TMEM -> registers, mask multiply, transposed
//     coalesced stores (consecutive lanes = consecutive i).
// Two CTAs per SM (88 KB of staging + 256 TMEM columns each) overlap one CTA's loads with the other's MMAs / epilogue.
#include "common.cuh"

using namespace ta;

namespace {

constexpr int kTileM = 128;          // stacked rows per CTA = TMEM lanes
constexpr int kBlockK = 32;          // K elements per staged block (8 core matrices of 4 tf32 each)
constexpr int kThreadsS = 256;
constexpr int kMaxN = 256;           // UMMA N limit; also the TMEM columns allocated

// shared-memory byte offset of element (row, kk) of a K-major no-swizzle operand block: 8-row groups of 1 KB, inside a group
// the 8 K-cores of 128 B (8 rows x 16 B) follow each other
__device__ __forceinline__ uint32_t core_off(int row, int kk) {
  return (uint32_t)((row >> 3) * 1024 + (kk >> 2) * 128 + (row & 7) * 16 + (kk & 3) * 4);
}

__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version = 1 [46,48),
  // layout type SWIZZLE_NONE = 0 [61,64)
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(128u >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// hi = v with the 13 low mantissa bits cleared (tf32 keeps 10); lo = v - hi (exact in fp32)
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  lo = sub_rn(v, hi);
}

// grid = ceil(planes*N / 128); dynamic smem: Ah | Al [128 x 32] + Wh | Wl [Npad x 32] fp32
template <bool SPLIT>
__global__ void __launch_bounds__(kThreadsS, 2) spectrum_gemm_kernel(const float* __restrict__ A, const float* __restrict__ addA,
                                                                      const float* __restrict__ W, const float* __restrict__ mulOut,
                                                                      float* __restrict__ out, int64_t rows_total, int N) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kTileM;
  const int n8 = (N + 7) & ~7;                                       // operand rows are staged in whole 8-row groups
  unsigned char* sAh = smem_raw;
  unsigned char* sAl = sAh + kTileM * kBlockK * 4;
  unsigned char* sWh = sAl + (SPLIT ? kTileM * kBlockK * 4 : 0);
  unsigned char* sWl = sWh + n8 * kBlockK * 4;

  if (warp == 0) {                                                   // TMEM: 256 columns (power of two >= N) for the fp32 accumulator
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"((uint32_t)kMaxN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(&s_bar, 1); mbar_fence_init(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;

  // InstrDescriptor (mma_sm100_desc.hpp): c_format F32 = 1 [4,6), a/b format TF32 = 2 [7,10) / [10,13), both K-major,
  // n_dim = N >> 3 [17,23), m_dim = 128 >> 4 [24,29)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

  const int nkb = (N + kBlockK - 1) / kBlockK;                       // K = N (square transforms)
  uint32_t phase = 0;
  for (int kb = 0; kb < nkb; ++kb) {
    const int k0 = kb * kBlockK;
    if (kb > 0) { mbar_wait(&s_bar, phase); phase ^= 1; }           // the MMAs reading the previous block have completed
    // ---- stage A block [128 x 32] (+ addA), split hi / lo ----
    for (int e = tid; e < kTileM * (kBlockK / 4); e += kThreadsS) {
      const int r = e >> 3, c4 = e & 7;                              // 8 float4 per row
      const int64_t gr = row0 + r;
      const int k = k0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < rows_total && k < N) {
        v = __ldg(reinterpret_cast<const float4*>(A + gr * N + k));
        if (addA) { const float4 w = __ldg(reinterpret_cast<const float4*>(addA + gr * N + k)); v.x = add_rn(v.x, w.x); v.y = add_rn(v.y, w.y); v.z = add_rn(v.z, w.z); v.w = add_rn(v.w, w.w); }
      }
      const uint32_t off = core_off(r, c4 * 4);
      if (SPLIT) {
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        *reinterpret_cast<float4*>(sAh + off) = h;
        *reinterpret_cast<float4*>(sAl + off) = l;
      } else {
        *reinterpret_cast<float4*>(sAh + off) = v;
      }
    }
    // ---- stage W block [N x 32] ----
    for (int e = tid; e < n8 * (kBlockK / 4); e += kThreadsS) {
      const int r = e >> 3, c4 = e & 7;
      const int k = k0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < N && k < N) v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)r * N + k));
      const uint32_t off = core_off(r, c4 * 4);
      if (SPLIT) {
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        *reinterpret_cast<float4*>(sWh + off) = h;
        *reinterpret_cast<float4*>(sWl + off) = l;
      } else {
        *reinterpret_cast<float4*>(sWh + off) = v;
      }
    }
    fence_proxy_async_smem();                                        // generic-proxy stores -> visible to the tensor core's async proxy
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      const uint32_t ah = smem_u32(sAh), al = smem_u32(sAl), wh = smem_u32(sWh), wl = smem_u32(sWl);
#pragma unroll
      for (int ks = 0; ks < kBlockK / 8; ++ks) {                     // one MMA covers K = 8 = two 16-byte cores = 256 B along K
        const uint32_t o = (uint32_t)ks * 256u;
        const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
        tcgen05_mma_tf32(tmem, umma_desc_kmajor(ah + o), umma_desc_kmajor(wh + o), idesc, first);
        if (SPLIT) {
          tcgen05_mma_tf32(tmem, umma_desc_kmajor(al + o), umma_desc_kmajor(wh + o), idesc, 1u);
          tcgen05_mma_tf32(tmem, umma_desc_kmajor(ah + o), umma_desc_kmajor(wl + o), idesc, 1u);
        }
      }
      tcgen05_commit(&s_bar);                                        // arrives when every MMA issued so far has completed
    }
  }
  mbar_wait(&s_bar, phase);                                          // accumulator complete
  tcgen05_fence_after();

  // ---- epilogue: TMEM -> registers -> out[plane][n][i] (* mulOut), 8 warps: lanes 32*(warp%4).., columns split by warp/4 ----
  {
    const int lane_base = (warp & 3) * 32;
    const int64_t gr = row0 + lane_base + lane;
    const bool live = gr < rows_total;
    const int64_t plane = live ? gr / N : 0;
    const int i = live ? (int)(gr - plane * N) : 0;
    float* obase = out + plane * (int64_t)N * N + i;
    const float* mbase = mulOut ? mulOut + plane * (int64_t)N * N + i : nullptr;
    const int half = warp >> 2;                                      // warps 0-3: column chunks 0,2,4..; warps 4-7: 1,3,5..
    for (int c0 = half * 16; c0 < N; c0 += 32) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)lane_base << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (live) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = c0 + j;
          if (n < N) {
            float v = __uint_as_float(r[j]);
            if (mbase) v = mul_rn(v, __ldg(mbase + (int64_t)n * N));
            obase[(int64_t)n * N] = v;
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kMaxN) : "memory");
}

template <bool SPLIT>
int launch_stage(const float* A, const float* addA, const float* W, const float* mulOut, float* out, int64_t rows_total, int N,
                 cudaStream_t s) {
  const int n8 = (N + 7) & ~7;
  size_t smem = (size_t)(SPLIT ? 2 : 1) * ((size_t)kTileM * kBlockK * 4 + (size_t)n8 * kBlockK * 4);
  if (smem < 80 * 1024) smem = 80 * 1024;       // never more than 2 CTAs per SM: each holds 256 of the SM's 512 TMEM columns
  auto k = spectrum_gemm_kernel<SPLIT>;
  static SmemOptIn optin = {};
  const int rc = ensure_dyn_smem("ta_spectrum_transform", k, smem, optin);
  if (rc != TA_OK) return rc;
  const int64_t blocks = (rows_total + kTileM - 1) / kTileM;
  TA_REQUIRE(blocks <= 0x7fffffff, "ta_spectrum_transform: too many rows");
  k<<<(unsigned)blocks, kThreadsS, smem, s>>>(A, addA, W, mulOut, out, rows_total, N);
  count_launch();
  return check_launch("ta_spectrum_transform");
}

}  // namespace

extern "C" {

int64_t ta_spectrum_ws_bytes(int planes, int N) { return (int64_t)2 * planes * N * N * (int64_t)sizeof(float); }

int ta_spectrum_transform(const float* x, const float* gauss, const float* mask, const float* D, const float* E, float* out,
                          int planes, int N, int precision, void* ws, ta_stream_t stream) {
  TA_REQUIRE(x && D && E && out && ws && planes > 0, "ta_spectrum_transform: null pointer or planes=%d", planes);
  if (N < 16 || N > kMaxN || N % 16 != 0) {
    set_error("ta_spectrum_transform: N=%d (needs a multiple of 16 in [16, %d])", N, kMaxN);
    return TA_EUNSUPPORTED;
  }
  TA_REQUIRE(aligned16(x) && aligned16(gauss) && aligned16(mask) && aligned16(D) && aligned16(E) && aligned16(out) && aligned16(ws),
             "ta_spectrum_transform: buffers must be 16-byte aligned");
  float* t0 = reinterpret_cast<float*>(ws);
  float* t1 = t0 + (int64_t)planes * N * N;
  const int64_t rows = (int64_t)planes * N;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
#define TA_STAGE(A_, ADD_, W_, MUL_, OUT_)                                                                         \
  rc = precision == 0 ? launch_stage<false>(A_, ADD_, W_, MUL_, OUT_, rows, N, s) : launch_stage<true>(A_, ADD_, W_, MUL_, OUT_, rows, N, s); \
  if (rc != TA_OK) return rc;
  TA_STAGE(x, gauss, D, nullptr, t0)          // R1 = D (x + gauss)^T
  TA_STAGE(t0, nullptr, D, mask, t1)          // Y  = (D X D^T) . mask
  TA_STAGE(t1, nullptr, E, nullptr, t0)       // R3 = E Y^T
  TA_STAGE(t0, nullptr, E, nullptr, out)      // T  = E Y E^T
#undef TA_STAGE
  return TA_OK;
}

}  // extern "C"
