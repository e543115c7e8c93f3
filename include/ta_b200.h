/*
 * ta_b200.h — C-ABI of libta_b200.so: the sm_100a kernels behind the TransferAttack
 * `Attack` hook API (reference: transferattack/attack.py and its gradient/,
 * input_transformation/, ensemble/ plugins).
 *
 * Conventions (every entry point):
 *   - plain C symbols, no torch / C++ types in any signature;
 *   - all tensor pointers are BORROWED device pointers to contiguous fp32 NCHW data
 *     (what `tensor.data_ptr()` returns); the caller keeps them alive until `stream`
 *     has been synchronised; nothing is allocated, freed or synchronised inside;
 *   - `stream` is a `cudaStream_t` passed as void* (0 = legacy default stream);
 *   - the return value is TA_OK (0) or a negative TA_E* code; the message for the last
 *     failure on the calling thread is returned by ta_last_error();
 *   - B = number of samples, n = elements per sample (C*H*W), N = total elements;
 *   - arithmetic is IEEE fp32, round-to-nearest, one rounding per reference op, never
 *     contracted into FMA unless the reference's own expression is an FMA (the
 *     bilinear source index, see ta_dim_fwd). This is what makes the results
 *     bit-comparable with the reference's eager PyTorch ops.
 *
 * Each declaration cites the reference code (file:line under the reference repo's
 * transferattack/ directory) that it replaces.
 */
#ifndef TA_B200_H
#define TA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TA_ABI_VERSION 1

enum {
  TA_OK = 0,
  TA_EINVAL = -1,       /* bad shape / null pointer / misaligned pointer */
  TA_ECUDA = -2,        /* CUDA runtime reported an error (launch, attribute, driver entry point) */
  TA_EUNSUPPORTED = -3  /* valid request this build cannot serve (e.g. kernel size too large) */
};

/* how ta_abs_mean_per_sample / ta_fused_update_linf form mean|g| */
enum {
  TA_MEAN_EXACT = 0,    /* fp64 accumulation, mean = (float)(sum / n): order-independent up to the final rounding. */
  TA_MEAN_TORCH = 1     /* the fp32 summation tree of torch's own CUDA kernel for `x.abs().mean(dim=(1,2,3))` (attack.py:128):
                           ATen's launch policy for the device (block shape, CTAs per output), 4 accumulators per thread,
                           shared-memory / shuffle trees, sum * (float)(B / numel) — the reference's bits without an ATen
                           launch. TA_EUNSUPPORTED outside the replayed launch family (B == 1, tiny samples): pass torch's own
                           result as `scale` there. */
};

/* direction modes of ta_update_linf */
enum {
  TA_DIR_SIGN = 0,      /* step = alpha * sign(dir)  (attack.py:147) */
  TA_DIR_RAW = 1        /* step = alpha * dir        (caller already holds a direction) */
};

typedef void* ta_stream_t;

/* ---- library ---------------------------------------------------------------------- */

int ta_version(void);                   /* TA_ABI_VERSION of the loaded library */
const char* ta_last_error(void);        /* thread-local message of the last non-OK return */
/* SM count / compute capability of the current device; TA_ECUDA when no device is usable. */
int ta_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of kernels this library has launched in this process since load (bench.py's gpu_launches) */
int64_t ta_launch_count(void);
/* runtime tuning knobs for the benchmark sweep ("fused.cluster", "fused.threads", "fused.unroll",
 * "fused.variant", "reduce.cluster", "dim.tma"); not part of the reference-facing surface. */
int ta_tune_set(const char* key, int value);

/* ---- get_momentum  (attack.py:124-128) ----------------------------------------------
 *   momentum * decay + grad / mean_{C,H,W}(|grad|)                                       */

/* mean_out[b] = mean over the sample of |g|.  mode: TA_MEAN_EXACT.
 * ws: caller-provided scratch of ta_abs_mean_ws_bytes(B, n) bytes (may be NULL when that is 0). */
int64_t ta_abs_mean_ws_bytes(int B, int64_t n);
int ta_abs_mean_per_sample(const float* g, float* mean_out, int B, int64_t n, int mode,
                           void* ws, ta_stream_t stream);
/* The launch policy TA_MEAN_TORCH replays (PyTorch ATen/native/cuda/Reduce.cuh setReduceConfig, fp32, vt0 = 4) for a device
 * with `sm_count` SMs and `max_threads_per_sm` resident threads per SM: block (block_w, block_h), ctas_per_output.
 * Host-only (no CUDA call). TA_EUNSUPPORTED when (B, n) is outside the replayed family.                                     */
int ta_aten_mean_policy(int B, int64_t n, int sm_count, int max_threads_per_sm, int* block_w, int* block_h,
                        int* ctas_per_output);

/* m_out = m * decay + g / scale[b]      (m == NULL means the reference's `momentum = 0` first call)
 * scale: [B] per-sample mean|g| (from torch or from ta_abs_mean_per_sample). m_out may alias m. */
int ta_momentum(const float* g, const float* m, const float* scale, float decay, float* m_out,
                int B, int64_t n, ta_stream_t stream);

/* ---- update_delta (attack.py:145-153), clamp (utils.py:68-69) --------------------------
 *   L-inf: delta' = clamp(clamp(delta + alpha * sign(dir), -eps, eps), lo - x, hi - x)
 *   alpha_t (nullable): full-shape [B,n] tensor step (gradient/gra.py:149, fgsra.py:213);
 *   when alpha_t != NULL the scalar `alpha` is ignored.  NaN propagates exactly as in
 *   torch.clamp / torch.min / torch.max.  delta_out may alias delta.                      */
int ta_update_linf(const float* delta, const float* data, const float* dir, const float* alpha_t,
                   float alpha, float eps, float lo, float hi, int dir_mode, float* delta_out,
                   int64_t N, ta_stream_t stream);

/* L2 (attack.py:148-152): ghat = g / (||g||_2 + 1e-20); y = delta + ghat * alpha;
 * rows with ||y||_2 > eps are scaled by eps / (||y||_2 + 1e-7) (torch.renorm); then box clamp.
 * ws: scratch of ta_update_l2_ws_bytes(B) bytes. */
int64_t ta_update_l2_ws_bytes(int B);
int ta_update_l2(const float* delta, const float* data, const float* g, float alpha, float eps,
                 float lo, float hi, float* delta_out, int B, int64_t n, void* ws, ta_stream_t stream);

/* init_delta's final projection (attack.py:141): out = min(max(delta, lo - x), hi - x). */
int ta_clamp_box(const float* delta, const float* data, float lo, float hi, float* out,
                 int64_t N, ta_stream_t stream);

/* L2 random start (attack.py:136-140): delta *= r / ||delta_row||_2 * eps, then box clamp.
 * delta holds the normal_ draw, r the uniform_(0,1) draw (both from torch's device generator). */
int ta_init_l2_scale(const float* delta, const float* r, const float* data, float eps, float lo, float hi,
                     float* out, int B, int64_t n, void* ws, ta_stream_t stream);

/* ---- the fused iteration tail (attack.py:97-100 + next iteration's attack.py:88) --------
 *   mu_b   = scale[b]                  if scale != NULL   (strict: torch computed it)
 *          = mean|g_b| (mean_mode)     otherwise          (fused in-kernel reduction)
 *   m'     = m * decay + g / mu_b      (m may be NULL on the first iteration: m' = 0 + g/mu_b)
 *   delta' = L-inf update_delta(delta, data, m', alpha)
 *   xadv   = data + delta'             (only when xadv_out != NULL; the next model input)
 * m_out / delta_out may alias m / delta (in place). scale_out (nullable, [B]) receives mu_b.
 * One launch; with the in-kernel reduction g is read from HBM once (cluster-resident).      */
int ta_fused_update_linf(const float* g, const float* m, float* m_out,
                         const float* delta, float* delta_out, const float* data,
                         float* xadv_out, const float* scale, float* scale_out, int mean_mode,
                         float decay, float alpha, float eps, float lo, float hi,
                         int B, int64_t n, ta_stream_t stream);

/* The same tail with every option, as one argument block (zero-initialise, then fill what applies):
 *   addend   (nullable) g' = g + addend before everything else — VMI/VNI's `grad + variance` (gradient/vmifgsm.py:87);
 *   gbar_out (nullable) receives g' / mu_b — EMI's bar_grad (gradient/emifgsm.py:97);
 *   emit_normalized / grad_wrt_xn + mean_host, std_host, C, plane: the Normalize fold of ta_fused_update_linf_nf below;
 *   delta_out != delta keeps the old delta intact (VMI evaluates its neighbours at the old point after the momentum update).
 * addend and grad_wrt_xn need n % 4 == 0, 16-byte aligned buffers and a sample that fits the cluster's shared memory
 * (TA_EUNSUPPORTED otherwise: keep the separate kernels).                                                                  */
typedef struct ta_fused_tail_args {
  const float* g; const float* addend;
  const float* m; float* m_out;
  const float* delta; float* delta_out;
  const float* data;
  float* xadv_out; float* gbar_out;
  const float* scale; float* scale_out;
  int mean_mode;
  float decay, alpha, eps, lo, hi;
  int B; int64_t n;
  const float* mean_host; const float* std_host; int C; int64_t plane; int emit_normalized; int grad_wrt_xn;
} ta_fused_tail_args;
int ta_fused_tail(const ta_fused_tail_args* args, ta_stream_t stream);

/* ---- Normalize folded into the fused tail (reference utils.py:72-79 PreprocessingModel) ------------------
 *   Same as ta_fused_update_linf, but the emitted next model input is the NORMALISED image
 *       xn = ((data + delta') - mean[c]) / std[c]        (torchvision Normalize: sub_ then div_, two roundings)
 *   with c = (element index inside the sample) / plane, so the surrogate is entered after its PreprocessingModel; with
 *   grad_wrt_xn != 0, `g` is the gradient w.r.t. xn and is first divided by std[c] (Normalize's adjoint) — then neither
 *   direction of the normalisation costs a launch. mean_host / std_host: HOST arrays [C] read during the call, C <= 4,
 *   plane % 4 == 0, C * plane == n, 16-byte aligned buffers; otherwise TA_EUNSUPPORTED (keep the separate kernels).      */
int ta_fused_update_linf_nf(const float* g, const float* m, float* m_out,
                            const float* delta, float* delta_out, const float* data, float* xn_out,
                            const float* scale, float* scale_out, int mean_mode,
                            float decay, float alpha, float eps, float lo, float hi, int B, int64_t n,
                            const float* mean_host, const float* std_host, int C, int64_t plane,
                            int grad_wrt_xn, ta_stream_t stream);

/* ---- ENS with one surrogate per GPU (ensemble/ens.py:31-36 + utils.py:94-100, new multi-GPU functionality) -----------
 *   The gradient reduce-scatter, the fused update and the all-gather of the next model input as ONE kernel over NVLink
 *   peer memory. Rank r owns samples [b0, b0+Bown). g_peers[k] / xadv_peers[k] (HOST arrays of K device pointers valid in
 *   this process: symmetric / IPC-mapped memory) are rank k's FULL [B, n] gradient and model-input buffers.
 *   For the owned samples: g = (((g_{K-1} + g_{K-2}) + ...) + g_0)  — the order autograd accumulates the members'
 *   gradients on one device — then ta_fused_update_linf's arithmetic; x_adv is stored into every rank's buffer, m' and
 *   delta' (full-batch pointers, only the owned rows are touched) stay local. The caller orders it across GPUs with a
 *   barrier before (all gradients written) and after (all x_adv visible) on the same stream. K <= 8.                 */
int ta_fused_allreduce_update_linf(const float* const* g_peers, float* const* xadv_peers, int K,
                                   const float* m, float* m_out, const float* delta, float* delta_out,
                                   const float* data, const float* scale, float* scale_out, int mean_mode,
                                   float decay, float alpha, float eps, float lo, float hi,
                                   int b0, int Bown, int64_t n, ta_stream_t stream);

/* ---- model-input staging (attack.py:88, gradient/nifgsm.py:35-39) -------------------------
 *   out = data + delta                         (look == NULL)
 *   out = (data + delta) + coef * look         (NI / VNI look-ahead; coef = alpha*decay as fp32)
 *   delta == NULL: `data` already holds the sum (out = data + coef * look).                     */
int ta_stage_add(const float* data, const float* delta, const float* look, float coef, float* out,
                 int64_t N, ta_stream_t stream);

/* PreprocessingModel's Normalize (utils.py:72-79): out = (x - mean[c]) / std[c]; adjoint gin = gout / std[c].
 * mean/std: [C] fp32 DEVICE arrays. plane = H*W. */
int ta_normalize_fwd(const float* x, const float* mean, const float* std, float* out,
                     int B, int C, int64_t plane, ta_stream_t stream);
int ta_normalize_bwd(const float* gout, const float* std, float* gin,
                     int B, int C, int64_t plane, ta_stream_t stream);

/* Normalize's adjoint (same bits as ta_normalize_bwd) that also leaves, per sample, the S column values of |gin| of torch's CUDA
 * `gin.abs().mean(dim=(1,2,3))` reduction (attack.py:128) in col_sums [B, S], S = block_w * block_h * ctas_per_output of
 * ta_aten_mean_policy for (B, C*plane) on the current device; ta_abs_mean_from_colsums finishes that mean (bit-identical to
 * torch's op, like TA_MEAN_TORCH) from those 4*S bytes per sample — the gradient is not read a second time. With mean_out [B]
 * and counters [B] (int32, zero before the first call; the kernel leaves them zero) the mean is finished inside the same launch:
 * every CTA reduces its block, the last CTA of a sample to arrive adds the per-block partials (col_sums is then only scratch for
 * those partials and does NOT hold the column values). Both NULL: column sums only.
 * C <= 4, plane % 4 == 0, 16-byte aligned tensors; TA_EUNSUPPORTED otherwise or outside the replayed launch family. */
int ta_normalize_bwd_colsum(const float* gout, const float* std, float* gin, float* col_sums,
                            float* mean_out, int* counters, int B, int C, int64_t plane, ta_stream_t stream);
int ta_abs_mean_from_colsums(const float* col_sums, float* mean_out, int B, int64_t n, ta_stream_t stream);

/* ---- SIM (input_transformation/sim.py:36-46) ----------------------------------------------
 *   out[s*N + i] = x[i] / 2^s, s = 0..S-1 (scale-major concat along the batch axis)
 *   adjoint: gin[i] = ((((g_{S-1}/2^{S-1}) + g_{S-2}/2^{S-2}) + ...) + g_0)  (autograd's accumulation order) */
int ta_sim_fwd(const float* x, float* out, int S, int64_t N, ta_stream_t stream);
int ta_sim_bwd(const float* gout, float* gin, int S, int64_t N, ta_stream_t stream);

/* ---- Admix (input_transformation/admix.py:40-51) --------------------------------------------
 *   out[((s*A + a)*B + b)] = (x[b] + strength * x[perm[a*B + b]]) / 2^s ; perm: int32 DEVICE array [A*B]
 *   (torch.randperm draws, made on the host generator by the caller).
 *   adjoint wrt the first x only (the mixed-in image is .detach()ed in the reference).           */
int ta_admix_fwd(const float* x, const int32_t* perm, float strength, float* out,
                 int S, int A, int B, int64_t n, ta_stream_t stream);
int ta_admix_bwd(const float* gout, float* gin, int S, int A, int B, int64_t n, ta_stream_t stream);

/* ---- DIM (input_transformation/dim.py:42-68) ---------------------------------------------------
 *   y1 = bilinear(x -> rnd x rnd); y2 = zero-pad y1 to R x R at (pad_top, pad_left);
 *   out = bilinear(y2 -> H x W); align_corners=False, no antialias; one (rnd, pad) for the batch.
 *   Source index = max(0, fmaf(scale, dst + 0.5f, -0.5f)), scale = (float)in / (float)out (ATen).
 *   planes = B*C; H == W required (the reference uses x.shape[-1] for both).
 *   ta_dim_bwd is the exact adjoint in deterministic gather form (ATen's uses atomicAdd).        */
int ta_dim_fwd(const float* x, float* out, int planes, int S, int rnd, int R, int pad_top, int pad_left,
               ta_stream_t stream);
int ta_dim_bwd(const float* gout, float* gin, int planes, int S, int rnd, int R, int pad_top, int pad_left,
               ta_stream_t stream);
/* The same two operations through the second-generation kernels, which keep their per-call tap / inverse-range tables in a
 * caller-provided DEVICE workspace of ta_dim_ws_bytes() bytes (16-byte aligned, alive until the stream has passed the call;
 * NULL falls back to the functions above). Forward: bit-identical to ta_dim_fwd; adjoint: same sums, other association.    */
int64_t ta_dim_ws_bytes(void);
int ta_dim_fwd_ws(const float* x, float* out, int planes, int S, int rnd, int R, int pad_top, int pad_left,
                  void* ws, ta_stream_t stream);
int ta_dim_bwd_ws(const float* gout, float* gin, int planes, int S, int rnd, int R, int pad_top, int pad_left,
                  void* ws, ta_stream_t stream);

/* DIM with the draw in DEVICE memory (for CUDA-graph replay; dim.py:47-62 draws a new (coin, rnd, pad_top, pad_left) per call):
 * `packs` is a device array of n_packs records of ta_dim_pack_bytes() bytes each, built on the HOST by ta_dim_pack_build (one per
 * pre-drawn iteration; identity != 0 = the coin said "return x") and uploaded by the caller; `it` is a device int32 holding the
 * index of the record to use (clamped to [0, n_packs-1]); ta_counter_add advances / resets it in stream order. The kernels are the
 * ones behind ta_dim_fwd_ws / ta_dim_bwd_ws reading geometry and tables from packs[*it] — bit-identical results. R = int(S*rate). */
int64_t ta_dim_pack_bytes(void);
int ta_dim_pack_build(void* host_pack, int S, int rnd, int R, int pad_top, int pad_left, int identity);
int ta_dim_fwd_dyn(const float* x, float* out, int planes, int S, int R, const void* packs, int n_packs,
                   const int* it, ta_stream_t stream);
int ta_dim_bwd_dyn(const float* gout, float* gin, int planes, int S, int R, const void* packs, int n_packs,
                   const int* it, ta_stream_t stream);
/* *counter = set_to >= 0 ? set_to : *counter + delta   (one-thread kernel, stream-ordered, capturable) */
int ta_counter_add(int* counter, int delta, int set_to, ta_stream_t stream);

/* ---- TIM (input_transformation/tim.py:68-73) ------------------------------------------------------
 *   out = conv2d(g, K[C,1,ks,ks], stride 1, zero padding 'same', groups=C)  (cross-correlation)
 *   k: [C, ks, ks] DEVICE array, ks odd, ks <= 31.
 *   ta_dwconv2d_sep: K[c] = outer(kcol[c], krow[c]) (rank-1 kernels: gaussian / uniform / linear of
 *   tim.py:42-66): out = sum_i kcol[i] * (sum_j krow[j] * g[y+i-r, x+j-r]).                          */
int ta_dwconv2d(const float* g, const float* k, int ks, float* out, int B, int C, int H, int W,
                ta_stream_t stream);
int ta_dwconv2d_sep(const float* g, const float* kcol, const float* krow, int ks, float* out,
                    int B, int C, int H, int W, ta_stream_t stream);
/* Same operation with the factors given as HOST arrays [C, ks] (read during the call): when all channels share them (every
 * kernel tim.py generates) and W % 4 == 0, 32 <= W <= 512, ks in {3,5,7,15}, they travel as kernel parameters and feed the
 * FMAs from the constant bank; any other request returns TA_EUNSUPPORTED (use ta_dwconv2d_sep). Bit-identical results.  */
int ta_dwconv2d_sep_hw(const float* g, const float* kcol_host, const float* krow_host, int ks,
                       float* out, int B, int C, int H, int W, ta_stream_t stream);

/* ---- PI-FGSM (gradient/pifgsm.py:55-68, 94-102) --------------------------------------------------------
 *   ta_pi_cut_noise:   amp' = amp + coef * sign(momentum)   (amp == NULL: the first iteration's python 0.0)
 *                      cut  = clamp(|amp'| - eps, 0, 10000) * sign(amp')
 *   project_noise (pifgsm.py:55-58) is ta_dwconv2d(cut, K) with K = ones/(k*k-1), centre 0.
 *   ta_pi_update_linf: proj = gamma * sign(conv);  amp'' = amp' + proj;
 *                      delta' = box(clamp((delta + alpha * sign(g)) + proj, -eps, eps))   (amp_out / delta_out may alias)   */
int ta_pi_cut_noise(const float* amp, const float* momentum, float coef, float eps, float* amp_out,
                    float* cut_out, int64_t N, ta_stream_t stream);
int ta_pi_update_linf(const float* delta, const float* data, const float* g, const float* conv,
                      const float* amp, float alpha, float gamma, float eps, float lo, float hi,
                      float* amp_out, float* delta_out, int64_t N, ta_stream_t stream);

/* ---- GRA / FGSRA decay indicator (gradient/gra.py:74-93, 148-149) -----------------------------------------
 *   eq = float(sign(last) == sign(cur));  M' = M * (eq + (1 - eq) * eta)         (last == NULL: the first iteration's python 0)
 *   delta' = L-inf update_delta(delta, data, cur, alpha_t = M' * alpha)           (attack.py:145-153 with a tensor step)
 *   One launch for the reference's 17 elementwise launches. M_out / delta_out may alias M / delta.                        */
int ta_gra_update(const float* M, const float* last, const float* cur, float eta, float alpha,
                  const float* delta, const float* data, float eps, float lo, float hi,
                  float* M_out, float* delta_out, int64_t N, ta_stream_t stream);

/* ---- AdaEA disparity-reduced filter (ensemble/adaea.py:115-136, 74-76, 82) ------------------------------------
 *   grads: HOST array of K (2..8) device pointers to the members' input gradients [B, C, plane], C in {1, 3}.
 *   Per pixel: u_k = normalize_C(g_k, eps 1e-12); cos(i,j) = cosine_similarity_C(u_i, u_j, eps 1e-8);
 *   r_i = (sum_{j != i} cos(i,j)) / (K-1) for i < K-1 (the reference's loop leaves the last member's row zero); map = mean_i r_i;
 *   mask = map >= threshold ? 1 : 0;  out = grad * mask.  map_out ([B, plane], nullable) receives the un-thresholded map;
 *   grad/out (nullable together) the filtered ensemble gradient. One launch for the reference's ~10 K^2 launches.        */
int ta_adaea_drf(const float* const* grads, int K, float threshold, const float* grad, float* out,
                 float* map_out, int B, int C, int64_t plane, ta_stream_t stream);

/* ---- SSM / FGSRA spectrum transform (input_transformation/ssm.py:41-55, 101-200) ----------------------------
 *   out = idct_2d(dct_2d(x + gauss) * mask) per [N x N] plane, with the reference's un-normalised DCT-II
 *   (X_k = 2 sum_n x_n cos(pi (2n+1) k / 2N)) and its exact inverse, evaluated as four tensor-core GEMMs (tcgen05, tf32
 *   operands, fp32 accumulation in TMEM):  T(X) = E ((D X D^T) . mask) E^T,  D[k][n] = 2 cos(pi (2n+1) k / 2N),  E = D^-1.
 *   D, E: [N, N] fp32 DEVICE matrices (row-major) supplied by the caller (built once per N); gauss / mask nullable;
 *   planes = B*C; N a multiple of 16 in [16, 256]. precision 1 (default) = "3xTF32": every operand is split hi + lo on the
 *   way into shared memory and hi*hi + lo*hi + hi*lo is accumulated (fp32-level products); precision 0 = one tf32 product.
 *   ws: scratch of ta_spectrum_ws_bytes(planes, N) bytes. 4 launches for the reference's ~40, no FFT.                      */
int64_t ta_spectrum_ws_bytes(int planes, int N);
int ta_spectrum_transform(const float* x, const float* gauss, const float* mask, const float* D, const float* E,
                          float* out, int planes, int N, int precision, void* ws, ta_stream_t stream);

/* ---- EMI (gradient/emifgsm.py:53-58, 86-103) ---------------------------------------------------------
 *   out[k*N + i] = x[i] + coef[k] * gbar[i]  (coef[k] = (float)(factor_k * alpha), host array, K <= 32)
 *   gbar == NULL is the first iteration (`bar_grad = 0`): out[k*N+i] = x[i] + 0.
 *   adjoint: gin[i] = (((g_{K-1}) + g_{K-2}) + ... ) + g_0                                              */
int ta_lin_sample_fwd(const float* x, const float* gbar, const float* coef_host, int K, float* out,
                      int64_t N, ta_stream_t stream);
int ta_lin_sample_bwd(const float* gout, float* gin, int K, int64_t N, ta_stream_t stream);

/* ---- VMI / VNI (gradient/vmifgsm.py:42-58) ---------------------------------------------------------------
 *   neighbour input: out = ((data + delta) + noise) [+ coef * look]   (noise = torch uniform_(-r, r) draw)
 *   accumulate:      acc = (first ? g : acc + g)        (`grad = 0; grad += ...`)
 *   finalize:        v = acc / num_neighbor - cur_grad                                                  */
int ta_neighbor_stage(const float* data, const float* delta, const float* noise, const float* look,
                      float coef, float* out, int64_t N, ta_stream_t stream);
/* The same staging with the noise generated in the kernel: noise[i] is bit for bit what torch's CUDA
 * `zeros_like(delta).uniform_(from, to)` (vmifgsm.py:50) would have written at element i for the device generator state
 * (seed, offset) — Philox4_32_10, torch's thread/element mapping (ATen DistributionTemplates.h) — so the attack consumes the
 * same random stream; the caller then advances the generator's offset by *offset_increment of ta_uniform_fill_policy(N).
 * One launch and 12 B/elem instead of 4 launches and 32 B/elem. noise_out (optional) receives the noise itself.
 * offset % 4 == 0, N < 2^31.                                                                                               */
int ta_uniform_fill_policy(int64_t numel, int64_t* threads_total, int64_t* offset_increment);
int ta_neighbor_stage_philox(const float* data, const float* delta, const float* look, float coef,
                             float from, float to, uint64_t seed, uint64_t offset,
                             float* out, float* noise_out, int64_t N, ta_stream_t stream);
int ta_accumulate(float* acc, const float* g, int first, int64_t N, ta_stream_t stream);
int ta_variance_finalize(const float* acc, const float* cur, int num_neighbor, float* out,
                         int64_t N, ta_stream_t stream);
/* out = a + b (vmifgsm.py:87 `grad + variance`) */
int ta_add(const float* a, const float* b, float* out, int64_t N, ta_stream_t stream);

/* ---- output path (utils.py:63-66 save_images) ---------------------------------------------------------------
 *   u8 = (uint8) trunc((data + delta) * 255)   as numpy's float32 -> uint8 cast of in-range values;
 *   layout NCHW -> NHWC (the permute((0,2,3,1)) of save_images) when to_nhwc != 0.                 */
int ta_quantize_u8(const float* data, const float* delta, uint8_t* out, int B, int C, int64_t plane,
                   int to_nhwc, ta_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TA_B200_H */
