// aten_mean.cuh — mean(|g|) per sample in the EXACT fp32 summation order of torch's CUDA `x.mean(dim=(1,2,3))`
// (reference call site: transferattack/attack.py:128 `grad.abs().mean(dim=(1,2,3), keepdim=True)`), so that the fused
// tail needs no ATen kernel and still produces the reference's bits (TA_MEAN_TORCH).
//
// What is replayed — PyTorch ATen/native/cuda/Reduce.cuh as shipped in the installed build's include tree (torch 2.11.0+cu128;
// restated in oracle/aten_reduce.py with line numbers; pinned against torch itself on the GPU box by tools/diag_aten_mean.py,
// tests/test_kernels_gpu.py and, at run time, by ops.aten_mean_replay_ok):
//   launch policy  setReduceConfig for a contiguous [B, n] fp32 tensor reduced over n, "vectorize along input": block (bw, bh),
//                  cpo CTAs per output; virtual thread t = tx + bw*ty + (bw*bh)*cta of an output owns the 128-bit vectors
//                  t, t+S, t+2S, ... with S = bw*bh*cpo;
//   thread_reduce  component i of each of its vectors is added (fp32, vectors in order) into accumulator i; value = ((a0+a1)+a2)+a3;
//   block_x_reduce FIRST: shared-memory tree over tx down to 32 lanes (offsets bw/2 .. 32), then shfl_down offsets 16,8,4,2,1;
//   block_y_reduce then the shared-memory tree over ty (offsets bh/2 .. 1);
//   global_reduce  the last CTA: partial i sits at linear thread id i (cpo <= bw: row ty = 0), y tree, then x tree;
//   MeanOps        mean = sum * factor, factor = (float)B / (float)(B*n).
//
// How it is mapped here: one thread-block cluster per sample. The sample is viewed as rows of S vectors; CTA r of the cluster
// owns the vector columns [r*W4, (r+1)*W4), W4 = S / cluster size — i.e. W4 of the S virtual threads, all their vectors.
// Phase 1: every vector column is reduced by one thread (the 4 accumulators are the 4 components) → its virtual thread's
// value in s_val[]. Phase 2 (after one cluster barrier): each CTA gathers all S values through DSMEM and replays the trees:
// a warp takes one (virtual block, ty) row at a time — its lanes hold tx = lane + 32k, halve in registers (the shared-memory
// levels) and shuffle (the warp levels) —, then one thread per virtual block does the y tree, then the final tree. Every CTA
// obtains the same mean; no global scratch, no atomics.
#pragma once

#include "common.cuh"

namespace ta {

struct AtenMeanCfg {
  int bw, bh, cpo;        // ATen's block shape and CTAs per output
  int nt;                 // bw * bh
  int S;                  // virtual threads per output = 128-bit vectors per row
  int W4;                 // vector columns (virtual threads) per CTA of the cluster
  unsigned long long w4_magic;   // floor(2^32 / W4) + 1: vt / W4 == (vt * magic) >> 32 for vt < S
  float factor;           // MeanOps factor
};

constexpr int kAtenThreads = 512;      // threads of the replay kernels (>= ATen's block size bw*bh)
constexpr int kAtenMaxW = 3584;        // s_val capacity (floats): covers cpo <= 56 with an 8-CTA cluster, cpo <= 28 with 4

// Host: ATen's launch policy. Returns false when the launch is outside the replayed family (n % 4 != 0 or n < 128: other
// load paths; small n: one warp row per output; cpo > bw).
bool aten_mean_policy(int B, int64_t n, int sm_count, int max_threads_per_sm, int* bw, int* bh, int* cpo);
// Host: policy for the current device + the cluster mapping; TA_OK / TA_EUNSUPPORTED (message set)
int aten_mean_plan(const char* who, int B, int64_t n, int cl, AtenMeanCfg* cfg);

// optional pre-processing of the gradient before |.|: Normalize's adjoint g / std_c (channel c = vector index / plane_vec) and
// an addend (VMI's grad + variance) — the same two steps, in the same order, as the fused tail applies them
struct MeanPre {
  const float* addend;
  float std[4];
  int64_t plane_vec;      // 128-bit vectors per channel plane (0: no division)
};
__device__ __forceinline__ float4 div4(float4 v, float s) { v.x = div_rn(v.x, s); v.y = div_rn(v.y, s); v.z = div_rn(v.z, s); v.w = div_rn(v.w, s); return v; }
__device__ __forceinline__ float4 add4(float4 a, const float4& b) { a.x = add_rn(a.x, b.x); a.y = add_rn(a.y, b.y); a.z = add_rn(a.z, b.z); a.w = add_rn(a.w, b.w); return a; }
__device__ __forceinline__ float pick4(const float (&a)[4], int c) { return c == 0 ? a[0] : (c == 1 ? a[1] : (c == 2 ? a[2] : a[3])); }
__device__ __forceinline__ int channel_of(int64_t vec, int64_t plane_vec) {
  return (vec >= plane_vec ? 1 : 0) + (vec >= 2 * plane_vec ? 1 : 0) + (vec >= 3 * plane_vec ? 1 : 0);
}

// ---- phase 1: one vector column ------------------------------------------------------------------------------------------
struct ColAcc { float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f; };
__device__ __forceinline__ void aten_column_add(ColAcc& A, const float4& v) {
  A.a0 = add_rn(A.a0, fabsf(v.x)); A.a1 = add_rn(A.a1, fabsf(v.y)); A.a2 = add_rn(A.a2, fabsf(v.z)); A.a3 = add_rn(A.a3, fabsf(v.w));
}
__device__ __forceinline__ float aten_column_value(const ColAcc& A) { return add_rn(add_rn(add_rn(A.a0, A.a1), A.a2), A.a3); }

// ---- phase 2: the trees -------------------------------------------------------------------------------------------------
// block_x_reduce of one block row held as a[k] = value[tx = lane + 32k], k < K = bw/32: lane 0 gets the sum
template <int K>
__device__ __forceinline__ float aten_x_tree(float (&a)[K]) {
#pragma unroll
  for (int h = K / 2; h >= 1; h >>= 1) {            // shared-memory levels: value[tx] += value[tx + 32h], tx < 32h
#pragma unroll
    for (int k = 0; k < h; ++k) a[k] = add_rn(a[k], a[k + h]);
  }
  float v = a[0];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = add_rn(v, __shfl_down_sync(0xffffffffu, v, o));   // warp levels, offsets decreasing
  return v;
}

// one warp reduces block rows warp, warp+16, ...: row = (virtual block, ty), bw = 32*K values each, gathered through DSMEM
// (virtual thread vt's value lives in CTA vt / W4 at s_val[vt % W4]). Two rows are in flight per step so that the remote
// loads of the second overlap the shuffles of the first.
// where virtual thread vt's column value lives: in the cluster's shared memory (CTA vt / W4 at s_val[vt % W4]) or in a global
// array of S values per sample (the column sums a previous kernel left: ta_normalize_bwd_colsum)
struct ColSrcCluster {
  const float* s_val; unsigned long long w4_magic; uint32_t W4;
  __device__ __forceinline__ float ld(uint32_t vt) const {
    const uint32_t owner = (uint32_t)(((unsigned long long)vt * w4_magic) >> 32);
    return dsmem_ld_f32(s_val + (vt - owner * W4), owner);
  }
};
struct ColSrcGlobal {
  const float* cs;
  __device__ __forceinline__ float ld(uint32_t vt) const { return __ldg(cs + vt); }
};
struct ColSrcShared {                                  // the sample's S values copied into this CTA's shared memory first
  const float* s;
  __device__ __forceinline__ float ld(uint32_t vt) const { return s[vt]; }
};

template <int K, class Src>
__device__ __forceinline__ void aten_rows_x_tree(const AtenMeanCfg& c, const Src& src, float* s_row, int warp, int lane) {
  const int nrows = c.cpo * c.bh;
  const int bw = 32 * K;
  auto fetch = [&](int row, float (&a)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = src.ld((uint32_t)(row * bw + lane + 32 * k));
  };
  int row = warp;
  for (; row + 16 < nrows; row += 32) {
    float a0[K], a1[K];
    fetch(row, a0); fetch(row + 16, a1);
    const float v0 = aten_x_tree<K>(a0), v1 = aten_x_tree<K>(a1);
    if (lane == 0) { s_row[row] = v0; s_row[row + 16] = v1; }
  }
  if (row < nrows) {
    float a0[K];
    fetch(row, a0);
    const float v0 = aten_x_tree<K>(a0);
    if (lane == 0) s_row[row] = v0;
  }
}

// s_val: this CTA's W4 column values (static shared memory, same offset in every CTA of the cluster), already written and
// made visible by a cluster barrier. s_row: >= cpo*bh floats, s_blk: >= max(cpo, 32) floats of CTA-local shared memory.
// Contains two __syncthreads(); all remote reads of s_val are complete after the first one. Returns the mean (same value in
// every thread of every CTA). blockDim.x == kAtenThreads.
template <class Src>
__device__ __forceinline__ float aten_tree_mean_src(const AtenMeanCfg& c, const Src& src, float* s_row, float* s_blk) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = c.bw >> 5;                              // 1 .. 16 values per lane before the shuffles
  switch (K) {
    case 1: aten_rows_x_tree<1>(c, src, s_row, warp, lane); break;
    case 2: aten_rows_x_tree<2>(c, src, s_row, warp, lane); break;
    case 4: aten_rows_x_tree<4>(c, src, s_row, warp, lane); break;
    case 8: aten_rows_x_tree<8>(c, src, s_row, warp, lane); break;
    default: aten_rows_x_tree<16>(c, src, s_row, warp, lane); break;
  }
  __syncthreads();
  if ((int)threadIdx.x < c.cpo) {                       // block_y_reduce of virtual block threadIdx.x (offsets bh/2 .. 1)
    const float* r = s_row + threadIdx.x * c.bh;
    float v;
    if (c.bh == 16) {
      float a[16];
#pragma unroll
      for (int y = 0; y < 16; ++y) a[y] = r[y];
#pragma unroll
      for (int h = 8; h >= 1; h >>= 1) {
#pragma unroll
        for (int y = 0; y < h; ++y) a[y] = add_rn(a[y], a[y + h]);
      }
      v = a[0];
    } else {
      float a[8];
#pragma unroll
      for (int y = 0; y < 8; ++y) a[y] = (y < c.bh) ? r[y] : 0.0f;
#pragma unroll
      for (int h = 4; h >= 1; h >>= 1)
        if (h < c.bh) {
#pragma unroll
          for (int y = 0; y < h; ++y) a[y] = add_rn(a[y], a[y + h]);
        }
      v = a[0];
    }
    s_blk[threadIdx.x] = v;
  }
  __syncthreads();
  // global_reduce's last block: partial i at (tx = i, ty = 0), identity elsewhere: the y tree adds +0.0f (exact); x tree.
  float v;
  if (c.cpo == 1) {
    v = s_blk[0];
  } else if (c.cpo <= 32) {                             // the shared-memory levels of the x tree only add +0.0f here
    float a[1] = {lane < c.cpo ? s_blk[lane] : 0.0f};
    v = aten_x_tree<1>(a);
  } else {
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { const int i = lane + 32 * k; a[k] = (k < K && i < c.cpo) ? s_blk[i] : 0.0f; }
    // K < 16: entries beyond K are zero, so the extra halving levels add +0.0f and the live levels are ATen's
    v = aten_x_tree<16>(a);
  }
  v = __shfl_sync(0xffffffffu, v, 0);
  return mul_rn(v, c.factor);
}

__device__ __forceinline__ float aten_tree_mean(const AtenMeanCfg& c, const float* s_val, float* s_row, float* s_blk) {
  return aten_tree_mean_src(c, ColSrcCluster{s_val, c.w4_magic, (uint32_t)c.W4}, s_row, s_blk);
}

// Normalize's adjoint that also leaves ATen's per-virtual-thread column sums of |gin| (ta_normalize_bwd_colsum), and the trees
// over such column sums (ta_abs_mean_from_colsums): the mean kernel's two halves, the first riding on a pass over the gradient
// that exists anyway
int aten_colsum_normalize_bwd(const float* gout, const float* std, float* gin, float* col_sums, float* mean_out, int* counters, int B, int C,
                              int64_t plane, cudaStream_t s);
int aten_colsum_tree(const float* col_sums, float* mean_out, int B, int64_t n, cudaStream_t s);

}  // namespace ta
