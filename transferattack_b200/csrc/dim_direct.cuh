// dim_direct.cuh — tables and entry points of the direct DIM kernels (dim_direct.cu), called from dim.cu's C-ABI functions.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ta {

constexpr int kDimMaxS = 512;     // image side the tables are sized for
constexpr int kDimMaxR = 576;     // >= int(kDimMaxS * 1.1) + a margin: padded size / intermediate size

constexpr int kDimRB = 16;        // destination rows per CTA (band height)

struct TapE { float l1; int i01; };          // 1-D bilinear tap of one destination index: i0 | i1 << 16, l0 = 1 - l1
struct InvE { short lo; short cnt; };        // destination indices [lo, lo + cnt) read this source index

struct DimTabF {                              // forward: 8.7 KB
  TapE t2[kDimMaxS];                          //   R -> S resize: output index o reads y2 indices
  TapE t1[kDimMaxR];                          //   S -> rnd resize: y1 index q reads source indices
};
struct DimTabB {                              // adjoint: 13 KB
  TapE t2[kDimMaxS];
  TapE t1[kDimMaxR];
  InvE inv2[kDimMaxR];                        //   y2 index p is read by outputs [lo, lo + cnt)
  InvE inv1[kDimMaxS];                        //   source index s is read by y1 indices [lo, lo + cnt)
  short4 band[kDimMaxS / kDimRB];             //   per band of source rows: {q0, nq, oyA, nu} (y1 rows / gout rows it needs)
};

bool dim_direct_ok(int S, int rnd, int R);
size_t dim_direct_ws_bytes();              // device workspace the direct kernels need for their tables (16-byte aligned)
int dim_fwd_direct(const float* x, float* out, int planes, int S, int rnd, int R, int top, int left, int blend, bool tma, void* ws,
                   cudaStream_t stream);
int dim_bwd_direct(const float* gout, float* gin, int planes, int S, int rnd, int R, int top, int left, bool tma, bool gather,
                   void* ws, cudaStream_t stream);

}  // namespace ta
