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
  float factor;           // MeanOps factor
};

constexpr int kAtenThreads = 512;      // threads of the replay kernels (>= ATen's block size bw*bh)
constexpr int kAtenMaxW = 3584;        // s_val capacity (floats): covers cpo <= 56 with an 8-CTA cluster, cpo <= 28 with 4

// Host: ATen's launch policy. Returns false when the launch is outside the replayed family (n % 4 != 0 or n < 128: other
// load paths; small n: one warp row per output; cpo > bw).
bool aten_mean_policy(int B, int64_t n, int sm_count, int max_threads_per_sm, int* bw, int* bh, int* cpo);
// Host: policy for the current device + the cluster mapping; TA_OK / TA_EUNSUPPORTED (message set)
int aten_mean_plan(const char* who, int B, int64_t n, int cl, AtenMeanCfg* cfg);

// ---- phase 1: one vector column ------------------------------------------------------------------------------------------
struct ColAcc { float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f; };
__device__ __forceinline__ void aten_column_add(ColAcc& A, const float4& v) {
  A.a0 = add_rn(A.a0, fabsf(v.x)); A.a1 = add_rn(A.a1, fabsf(v.y)); A.a2 = add_rn(A.a2, fabsf(v.z)); A.a3 = add_rn(A.a3, fabsf(v.w));
}
__device__ __forceinline__ float aten_column_value(const ColAcc& A) { return add_rn(add_rn(add_rn(A.a0, A.a1), A.a2), A.a3); }

// ---- phase 2: the trees -------------------------------------------------------------------------------------------------
// block_x_reduce of one block row held as a[k] = value[tx = lane + 32k], k < K = bw/32 (zero beyond): lane 0 gets the sum
template <int KMAX>
__device__ __forceinline__ float aten_x_tree(float (&a)[KMAX], int K) {
#pragma unroll
  for (int h = KMAX / 2; h >= 1; h >>= 1)            // shared-memory levels: value[tx] += value[tx + 32h], tx < 32h
    if (h < K) {
#pragma unroll
      for (int k = 0; k < KMAX / 2; ++k) if (k < h) a[k] = add_rn(a[k], a[k + h]);
    }
  float v = a[0];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = add_rn(v, __shfl_down_sync(0xffffffffu, v, o));   // warp levels, offsets decreasing
  return v;
}

// s_val: this CTA's W4 column values (static shared memory, same offset in every CTA of the cluster), already written and
// made visible by a cluster barrier. s_row: >= cpo*bh floats, s_blk: >= bw floats of CTA-local shared memory.
// Contains two __syncthreads(); all remote reads of s_val are complete after the first one. Returns the mean (same value in
// every thread of every CTA). blockDim.x == kAtenThreads.
__device__ __forceinline__ float aten_tree_mean(const AtenMeanCfg& c, const float* s_val, float* s_row, float* s_blk) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = c.bw >> 5;                              // 1 .. 16 values per lane before the shuffles
  const int nrows = c.cpo * c.bh;                       // (virtual block, ty) rows of bw values each
  for (int row = warp; row < nrows; row += kAtenThreads / 32) {
    const int base = row * c.bw;                        // = cb*nt + ty*bw : first virtual thread of the row
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      a[k] = 0.0f;
      if (k < K) {
        const int vt = base + lane + 32 * k;
        const int owner = vt / c.W4;
        a[k] = dsmem_ld_f32(s_val + (vt - owner * c.W4), (uint32_t)owner);
      }
    }
    const float v = aten_x_tree<16>(a, K);
    if (lane == 0) s_row[row] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < c.cpo) {                       // block_y_reduce of virtual block threadIdx.x
    float a[16];
#pragma unroll
    for (int y = 0; y < 16; ++y) a[y] = (y < c.bh) ? s_row[threadIdx.x * c.bh + y] : 0.0f;
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1)
      if (h < c.bh) {
#pragma unroll
        for (int y = 0; y < 8; ++y) if (y < h) a[y] = add_rn(a[y], a[y + h]);
      }
    s_blk[threadIdx.x] = a[0];
  }
  __syncthreads();
  // global_reduce's last block: partial i at (tx = i, ty = 0), identity elsewhere: the y tree adds +0.0f (exact); x tree.
  float v;
  if (c.cpo > 1) {
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { const int i = lane + 32 * k; a[k] = (k < K && i < c.cpo) ? s_blk[i] : 0.0f; }
    v = aten_x_tree<16>(a, K);
  } else {
    v = s_blk[0];
  }
  v = __shfl_sync(0xffffffffu, v, 0);
  return mul_rn(v, c.factor);
}

}  // namespace ta
