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

// One draw of DIM (dim.py:47-62: coin, rnd, pad_top, pad_left) with everything the direct kernels derive from it, as plain data:
// an array of these in DEVICE memory plus a device iteration counter lets ONE captured CUDA graph serve every iteration's draw.
struct DimGeo { int S, rnd, R, top, left, a_rows, c_rows, pad; };
struct DimPack {
  int identity;            // the coin said "do not transform": both directions copy
  int pad[3];
  DimGeo gf, gb;           // forward / adjoint geometry (band row maxima of THIS draw)
  DimGeo gw;               // forward, source-driven walk (32-row bands): a_rows = source rows, pad = y2 rows per band (0: not applicable)
  DimTabF tf;
  DimTabB tb;
};

bool dim_direct_ok(int S, int rnd, int R);
// host: fill `pack` for one draw; shared-memory bytes per CTA that serve EVERY possible draw at (S, R)
int dim_pack_build(DimPack* pack, int S, int rnd, int R, int top, int left, int identity);
void dim_dyn_smem(int S, int R, size_t* fwd_bytes, size_t* bwd_bytes, size_t* fwd_sep_bytes = nullptr, size_t* bwd_sep_bytes = nullptr,
                  size_t* fwd_walk_bytes = nullptr);
int dim_fwd_dyn(const float* x, float* out, int planes, int S, int R, const DimPack* packs, int n_packs, const int* it, bool tma,
                cudaStream_t stream);
int dim_bwd_dyn(const float* gout, float* gin, int planes, int S, int R, const DimPack* packs, int n_packs, const int* it, bool tma,
                cudaStream_t stream);
size_t dim_direct_ws_bytes();              // device workspace the direct kernels need for their tables (16-byte aligned)
int dim_fwd_direct(const float* x, float* out, int planes, int S, int rnd, int R, int top, int left, int blend, bool tma, void* ws,
                   cudaStream_t stream);
int dim_bwd_direct(const float* gout, float* gin, int planes, int S, int rnd, int R, int top, int left, bool tma, bool gather,
                   void* ws, cudaStream_t stream);

}  // namespace ta
