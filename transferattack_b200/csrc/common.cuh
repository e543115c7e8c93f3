// common.cuh — shared device/host helpers for libta_b200.so (sm_100a only).
//
// Arithmetic contract: IEEE fp32, round-to-nearest-even, ONE rounding per
// reference op. Every arithmetic step that must not be contracted goes through the __f*_rn
// intrinsics (never fused by ptxas, independent of -fmad). The library is additionally built with
// -fmad=false so that incidental expressions are not contracted either.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ta_b200.h"

namespace ta {

// ---- host-side error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);          // cudaGetLastError() → TA_OK / TA_ECUDA (+message)
void count_launch(int n = 1);
int sm_count();                              // cached multiProcessorCount of the current device
int tune_get(const char* key, int dflt);     // runtime tuning knobs (ta_tune_set)

#define TA_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      ta::set_error(__VA_ARGS__);             \
      return TA_EINVAL;                       \
    }                                         \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- exact elementwise arithmetic -----------------------------------------------------------------
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

// torch.sign: (0 < x) - (x < 0)  → sign(NaN) = 0, sign(±0) = +0
__device__ __forceinline__ float sign_t(float v) { return (float)((0.0f < v) - (v < 0.0f)); }
// torch.max / torch.min / torch.clamp: NaN in either operand propagates (fmaxf/fminf would drop it)
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ float min_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

// attack.py:147 + :152 / utils.py:68-69 — one element of the L-inf projection
__device__ __forceinline__ float project_linf(float delta, float step, float x, float eps, float lo, float hi) {
  const float d1 = add_rn(delta, step);
  const float d2 = min_nan(max_nan(d1, -eps), eps);
  return min_nan(max_nan(d2, sub_rn(lo, x)), sub_rn(hi, x));
}

// ---- small fixed-size vectors (V = 1 scalar fallback, V = 4 → one 128-bit access) --------------------
template <int V> struct Vec { float v[V]; };

template <int V> __device__ __forceinline__ Vec<V> ldv(const float* __restrict__ p, int64_t i);
template <> __device__ __forceinline__ Vec<1> ldv<1>(const float* __restrict__ p, int64_t i) {
  Vec<1> r; r.v[0] = __ldg(p + i); return r;
}
template <> __device__ __forceinline__ Vec<4> ldv<4>(const float* __restrict__ p, int64_t i) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
  Vec<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
// plain (coherent) loads for buffers that may alias an output of the same kernel (in-place updates)
template <int V> __device__ __forceinline__ Vec<V> ldv_rw(const float* p, int64_t i);
template <> __device__ __forceinline__ Vec<1> ldv_rw<1>(const float* p, int64_t i) { Vec<1> r; r.v[0] = p[i]; return r; }
template <> __device__ __forceinline__ Vec<4> ldv_rw<4>(const float* p, int64_t i) {
  const float4 t = reinterpret_cast<const float4*>(p)[i];
  Vec<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
template <int V> __device__ __forceinline__ void stv(float* p, int64_t i, const Vec<V>& a);
template <> __device__ __forceinline__ void stv<1>(float* p, int64_t i, const Vec<1>& a) { p[i] = a.v[0]; }
template <> __device__ __forceinline__ void stv<4>(float* p, int64_t i, const Vec<4>& a) {
  reinterpret_cast<float4*>(p)[i] = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}

// ---- warp / block reductions (fixed order → deterministic) ---------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// all threads get the block total; `scratch` must hold blockDim.x/32 doubles; contains 2 __syncthreads
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += scratch[i];   // same order in every thread
  __syncthreads();
  return t;
}

// ---- mbarrier + bulk-TMA (cp.async.bulk, SASS UBLKCP) PTX wrappers ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded spin: a lost completion traps (kernel error) instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
// global → this CTA's shared memory, completion signalled on `bar` (bytes % 16 == 0, both 16-B aligned)
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared → global bulk store (bulk_group completion)
__device__ __forceinline__ void tma_bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes → visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- thread-block-cluster helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }
// read a double from the same smem offset in CTA `rank` of this cluster (DSMEM)
__device__ __forceinline__ double dsmem_ld_f64(const double* local, uint32_t rank) {
  uint32_t remote; double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
  return v;
}
__device__ __forceinline__ float dsmem_ld_f32(const float* local, uint32_t rank) {
  uint32_t remote; float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

// ---- generic vectorised elementwise launchers ------------------------------------------------------------------------
// Functors come in two shapes:
//   (a) two-phase (the hot kernels):  template <int V> __device__ L load(i) const;  template <int V> __device__ void apply(i, const L&) const;
//       the kernel issues the loads of U vectors per thread before the first use (U x #inputs 128-bit loads in flight per
//       thread — what a streaming kernel needs on HBM3e), then computes and stores. Outputs may alias inputs (same index).
//   (b) single-phase:                 template <int V> __device__ void run(i) const;      wrapped by OnePhase<F>, U = 1.
// The grid covers the whole range in one pass (no cap, like ATen's elementwise launches): one batch of U vectors per thread.
template <class F> struct OnePhase {
  F f;
  template <int V> __device__ __forceinline__ int load(int64_t) const { return 0; }
  template <int V> __device__ __forceinline__ void apply(int64_t i, int) const { f.template run<V>(i); }
  template <int V> __device__ __forceinline__ int load(int, int64_t) const { return 0; }
  template <int V> __device__ __forceinline__ void apply(int row, int64_t i, int) const { f.template run<V>(row, i); }
};

template <int V, int U, class F>
__global__ void __launch_bounds__(256) ew_kernel(int64_t nvec, F f) {
  // each CTA owns U * blockDim consecutive vectors per step (contiguous 4 KB pieces per stream: DRAM-page friendly)
  const int64_t step = (int64_t)gridDim.x * blockDim.x * U;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x * U + threadIdx.x; i0 < nvec; i0 += step) {
    decltype(f.template load<V>(i0)) ld[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int64_t i = i0 + u * (int64_t)blockDim.x; if (i < nvec) ld[u] = f.template load<V>(i); }
#pragma unroll
    for (int u = 0; u < U; ++u) { const int64_t i = i0 + u * (int64_t)blockDim.x; if (i < nvec) f.template apply<V>(i, ld[u]); }
  }
}

template <int U, class F>
int launch_ew2(const char* name, int64_t N, bool can_vec4, F f, cudaStream_t s) {
  if (N <= 0) return TA_OK;
  const int threads = 256;
  const int64_t nvec = can_vec4 ? N / 4 : N;
  int64_t want = (nvec + (int64_t)threads * U - 1) / ((int64_t)threads * U);
  if (want > 0x7fffffff) want = 0x7fffffff;
  if (can_vec4) ew_kernel<4, U, F><<<(unsigned)want, threads, 0, s>>>(nvec, f);
  else ew_kernel<1, U, F><<<(unsigned)want, threads, 0, s>>>(nvec, f);
  count_launch();
  return check_launch(name);
}
template <class F>
int launch_ew(const char* name, int64_t N, bool can_vec4, F f, cudaStream_t s) {
  return launch_ew2<1>(name, N, can_vec4, OnePhase<F>{f}, s);
}

// per-row variant: blockIdx.y = row (a sample, or a (sample, channel) plane); functors get (row, i) with i the V-wide vector
// index inside the row, so per-row constants (scale[b], mean[c]) cost no 64-bit division per vector
template <int V, int U, class F>
__global__ void __launch_bounds__(256) ew_rows_kernel(int64_t nvec_per_row, F f) {
  const int row = blockIdx.y;
  const int64_t step = (int64_t)gridDim.x * blockDim.x * U;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x * U + threadIdx.x; i0 < nvec_per_row; i0 += step) {
    decltype(f.template load<V>(row, i0)) ld[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int64_t i = i0 + u * (int64_t)blockDim.x; if (i < nvec_per_row) ld[u] = f.template load<V>(row, i); }
#pragma unroll
    for (int u = 0; u < U; ++u) { const int64_t i = i0 + u * (int64_t)blockDim.x; if (i < nvec_per_row) f.template apply<V>(row, i, ld[u]); }
  }
}

// cap_per_sm > 0: persistent launch — at most cap_per_sm resident CTAs per SM over all rows, each looping over its row
template <int U, class F>
int launch_ew_rows2(const char* name, int rows, int64_t n_per_row, bool can_vec4, F f, cudaStream_t s, int cap_per_sm = 0) {
  if (rows <= 0 || n_per_row <= 0) return TA_OK;
  if (rows > 65535) { set_error("%s: %d rows exceed the grid limit 65535", name, rows); return TA_EINVAL; }
  const int threads = 256;
  const int64_t nvec = can_vec4 ? n_per_row / 4 : n_per_row;
  int64_t want = (nvec + (int64_t)threads * U - 1) / ((int64_t)threads * U);
  if (want > 0x7fffffff) want = 0x7fffffff;
  if (cap_per_sm > 0) {
    int64_t per_row = ((int64_t)sm_count() * cap_per_sm + rows - 1) / rows;
    if (per_row < 1) per_row = 1;
    if (want > per_row) want = per_row;
  }
  const dim3 grid((unsigned)want, (unsigned)rows);
  if (can_vec4) ew_rows_kernel<4, U, F><<<grid, threads, 0, s>>>(nvec, f);
  else ew_rows_kernel<1, U, F><<<grid, threads, 0, s>>>(nvec, f);
  count_launch();
  return check_launch(name);
}
template <class F>
int launch_ew_rows(const char* name, int rows, int64_t n_per_row, bool can_vec4, F f, cudaStream_t s) {
  return launch_ew_rows2<1>(name, rows, n_per_row, can_vec4, OnePhase<F>{f}, s);
}

// ---- cluster-wide sum + cluster launch --------------------------------------------------------------------------
// Sum over the cluster of a per-thread double; every thread of every CTA receives the same total (combined in
// rank order → deterministic). s_scratch: >= 32 doubles, s_part: 1 double (both CTA-local shared memory).
__device__ __forceinline__ double cluster_allreduce_sum(double v, double* s_scratch, double* s_part) {
  const double t = block_sum(v, s_scratch);
  if (threadIdx.x == 0) *s_part = t;
  cluster_sync_all();
  double tot = 0.0;
  const uint32_t nr = cluster_nctarank();
  for (uint32_t r = 0; r < nr; ++r) tot += dsmem_ld_f64(s_part, r);
  cluster_sync_all();   // nobody may overwrite / retire s_part before all ranks have read it
  return tot;
}

int pick_cluster(int64_t n, int threads);

// Opt a kernel in to `bytes` of dynamic shared memory (> 48 KB needs it) once per (kernel, device): the attribute call is
// skipped when an equal or larger size was granted before, so steady-state launches (and CUDA-graph captures) issue none.
struct SmemOptIn { size_t granted[64]; };
template <class K>
int ensure_dyn_smem(const char* who, K kernel, size_t bytes, SmemOptIn& st) {
  if (bytes <= 48 * 1024) return TA_OK;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (st.granted[dev] >= bytes) return TA_OK;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cannot reserve %zu B of shared memory: %s", who, bytes, cudaGetErrorString(e));
    cudaGetLastError();
    return TA_ECUDA;
  }
  st.granted[dev] = bytes;
  return TA_OK;
}   // CTAs per sample for the per-sample reduction kernels (<= 8)

template <class T> struct ident { using type = T; };

// grid = (cluster, B): one cluster of `cl` CTAs per sample, blockIdx.y = sample
template <class... Args>
int launch_cluster(const char* name, void (*kernel)(Args...), int cl, int B, int threads, size_t smem, cudaStream_t s,
                   typename ident<Args>::type... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cl, (unsigned)B, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  count_launch();
  if (e != cudaSuccess) {
    set_error("%s: cudaLaunchKernelEx failed: %d (%s)", name, (int)e, cudaGetErrorString(e));
    cudaGetLastError();
    return TA_ECUDA;
  }
  return check_launch(name);
}

}  // namespace ta
