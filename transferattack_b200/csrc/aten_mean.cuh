// aten_mean.cuh — mean(|g|) per sample in the EXACT fp32 summation order of torch's CUDA `x.mean(dim=(1,2,3))`
// (reference call site: transferattack/attack.py:128 `grad.abs().mean(dim=(1,2,3), keepdim=True)`), so that the fused
// tail needs no ATen kernel and still produces the reference's bits (TA_MEAN_TORCH).
//
// What is replayed (PyTorch ATen/native/cuda/Reduce.cuh, restated in oracle/aten_reduce.py and pinned against torch itself
// on the GPU box by tools/diag_aten_mean.py and tests/test_kernels_gpu.py):
//   launch policy  setReduceConfig for a contiguous [B, n] fp32 tensor reduced over n: block (bw, bh), bw*bh = 512,
//                  cpo CTAs per output; "virtual thread" t = tx + bw*ty + 512*cta of an output owns elements t, t+S, t+2S, ...
//                  with S = 512*cpo;
//   thread_reduce  element j of a virtual thread is added (fp32, in order) into accumulator j % 4; value = ((a0+a1)+a2)+a3;
//   block_y_reduce shared-memory tree over ty (offsets bh/2 .. 1);
//   block_x_reduce shared-memory tree over tx down to 32 lanes (offsets bw/2 .. 32), then shfl_down offsets 1,2,4,8,16;
//   global_reduce  the last CTA sums the cpo partials with the same two trees (partial i sits at tx = i, ty = 0);
//   MeanOps        mean = sum * factor, factor = (float)B / (float)(B*n).
//
// How it is mapped here: one thread-block cluster per sample. The sample is viewed as rows of S elements; CTA r of the
// cluster owns the columns [r*W, (r+1)*W), W = S / cluster size — i.e. W of the S virtual threads, all their elements.
// Phase 1: every column is reduced by one thread (4 accumulators, rows in order) → its virtual thread's value in s_val[].
// Phase 2 (after one cluster barrier): each CTA gathers all S values through DSMEM, one per (virtual block, thread
// position), and replays the trees with warp shuffles: the y tree inside bh-lane groups, the x tree after one
// shared-memory transpose. Every CTA obtains the same mean; no global scratch, no atomics.
#pragma once

#include "common.cuh"

namespace ta {

struct AtenMeanCfg {
  int bw, bh, cpo;        // ATen's block shape and CTAs per output
  int S;                  // virtual threads per output = elements per row
  int W;                  // columns (virtual threads) per CTA of the cluster
  float factor;           // MeanOps factor
};

constexpr int kAtenThreads = 512;      // ATen's block size for 4-byte types: the replay needs exactly this many threads
constexpr int kAtenMaxW = 2560;        // s_val capacity (floats): covers cpo <= 20 with a 4-CTA cluster, cpo <= 40 with 8

// Host: ATen's launch policy. Returns false when the launch is outside the replayed family (B == 1: vectorised 1-D path;
// small n: one warp row per output; cpo beyond what the final tree here covers).
bool aten_mean_policy(int B, int64_t n, int sm_count, int max_threads_per_sm, int* bw, int* bh, int* cpo);
// Host: policy for the current device + the cluster mapping; TA_OK / TA_EUNSUPPORTED (message set)
int aten_mean_plan(const char* who, int B, int64_t n, int cl, AtenMeanCfg* cfg);

// dynamic shared memory (floats) phase 2 needs for the transposed y-tree results: cpo * bw
__host__ __device__ inline int aten_mean_tree_floats(const AtenMeanCfg& c) { return c.cpo * c.bw; }

// ---- phase 1: one column -------------------------------------------------------------------------------------------
// thread_reduce_impl's order: element j of the column is added (fp32) into accumulator j % 4, rows in order; the value is
// ((a0+a1)+a2)+a3. fetch(j) returns |element| of row j. The rows may be fed in segments [j0, j1) whose starts are multiples
// of 4 (the fused kernel waits for one transfer group per segment).
struct ColAcc { float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f; };
template <class Fetch>
__device__ __forceinline__ void aten_column_rows(ColAcc& A, int j0, int j1, Fetch fetch) {
  int j = j0;
  for (; j + 3 < j1; j += 4) {
    const float v0 = fetch(j), v1 = fetch(j + 1), v2 = fetch(j + 2), v3 = fetch(j + 3);
    A.a0 = add_rn(A.a0, v0); A.a1 = add_rn(A.a1, v1); A.a2 = add_rn(A.a2, v2); A.a3 = add_rn(A.a3, v3);
  }
  if (j < j1) A.a0 = add_rn(A.a0, fetch(j));
  if (j + 1 < j1) A.a1 = add_rn(A.a1, fetch(j + 1));
  if (j + 2 < j1) A.a2 = add_rn(A.a2, fetch(j + 2));
}
__device__ __forceinline__ float aten_column_value(const ColAcc& A) { return add_rn(add_rn(add_rn(A.a0, A.a1), A.a2), A.a3); }

// ---- phase 2: the trees -------------------------------------------------------------------------------------------------
// s_val: this CTA's W column values (static shared memory, same offset in every CTA of the cluster), already written and
// made visible by a cluster barrier. s_tree: >= cpo*bw floats of CTA-local shared memory. s_blk: >= 32 floats.
// Contains two __syncthreads(); remote reads of s_val are complete after the first one (the caller may then arrive on the
// cluster barrier that protects s_val). Returns the mean (same value in every thread of every CTA). blockDim.x == 512.
__device__ __forceinline__ float aten_tree_mean(const AtenMeanCfg& c, const float* s_val, float* s_tree, float* s_blk) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ty = lane & (c.bh - 1);                 // bh is a power of two <= 16
  const int tx = warp * (32 / c.bh) + lane / c.bh;  // 16 warps * (32/bh) = bw
  const int pos = ty * c.bw + tx;
  for (int cb = 0; cb < c.cpo; ++cb) {
    const int vt = cb * kAtenThreads + pos;
    const int owner = vt / c.W;
    float v = dsmem_ld_f32(s_val + (vt - owner * c.W), (uint32_t)owner);
    for (int h = c.bh >> 1; h >= 1; h >>= 1) v = add_rn(v, __shfl_down_sync(0xffffffffu, v, h, c.bh));   // block_y_reduce
    if (ty == 0) s_tree[cb * c.bw + tx] = v;
  }
  __syncthreads();
  const int K = c.bw >> 5;                           // 1, 2, 4 or 8 values per lane before the shuffles
  for (int cb = warp; cb < c.cpo; cb += kAtenThreads / 32) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (k < K) ? s_tree[cb * c.bw + lane + 32 * k] : 0.0f;
#pragma unroll
    for (int h = 4; h >= 1; h >>= 1)                 // block_x_reduce, shared-memory levels: value[tx] += value[tx + 32h]
      if (h < K) {
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k < h) a[k] = add_rn(a[k], a[k + h]);
      }
    float v = a[0];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v = add_rn(v, __shfl_down_sync(0xffffffffu, v, o));                  // warp levels
    if (lane == 0) s_blk[cb] = v;
  }
  __syncthreads();
  // global_reduce's last block: partial i at (tx = i, ty = 0), identity elsewhere; cpo <= 32 <= bw, so the y tree and the
  // shared-memory x levels only add +0.0f (exact) and the warp levels decide. Every warp does it: no broadcast needed.
  float v = (lane < c.cpo) ? s_blk[lane] : 0.0f;
  if (c.cpo > 1) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v = add_rn(v, __shfl_down_sync(0xffffffffu, v, o));
  }
  v = __shfl_sync(0xffffffffu, v, 0);
  return mul_rn(v, c.factor);
}

}  // namespace ta
