// reduce.cu — per-sample reductions: mean|g| (attack.py:128), the L2 update (attack.py:148-152) and the L2
// random start (attack.py:136-140). One thread-block CLUSTER per sample: every CTA reduces a slice, partials
// are combined through distributed shared memory in rank order (deterministic, no atomics, no scratch).
#include "common.cuh"

using namespace ta;

namespace ta { struct MeanPre; int aten_abs_mean_launch(const float* g, float* mean_out, int B, int64_t n, const MeanPre* pre, cudaStream_t s); }

namespace {

constexpr int kThreads = 512;

struct Slice { int64_t begin, end; };   // in units of V-wide vectors, relative to the sample
__device__ __forceinline__ Slice my_slice(int64_t nvec) {
  const int64_t nr = cluster_nctarank(), r = cluster_ctarank();
  const int64_t per = (nvec + nr - 1) / nr;
  Slice s;
  s.begin = r * per < nvec ? r * per : nvec;
  s.end = (r + 1) * per < nvec ? (r + 1) * per : nvec;
  return s;
}

// grid = (cluster, B): blockIdx.y = sample
template <int V>
__global__ void __launch_bounds__(kThreads) abs_mean_kernel(const float* __restrict__ g, float* __restrict__ mean_out, int64_t n) {
  __shared__ double s_scratch[32];
  __shared__ double s_part;
  const int64_t nvec = n / V;
  const float* gp = g + (int64_t)blockIdx.y * n;
  const Slice sl = my_slice(nvec);
  double acc = 0.0;
#pragma unroll 4
  for (int64_t i = sl.begin + threadIdx.x; i < sl.end; i += kThreads) {
    const Vec<V> v = ldv<V>(gp, i);
#pragma unroll
    for (int k = 0; k < V; ++k) acc += (double)fabsf(v.v[k]);
  }
  const double tot = cluster_allreduce_sum(acc, s_scratch, &s_part);
  if (cluster_ctarank() == 0 && threadIdx.x == 0) mean_out[blockIdx.y] = (float)(tot / (double)n);
}

// attack.py:148-152
template <int V>
__global__ void __launch_bounds__(kThreads) update_l2_kernel(const float* delta, const float* __restrict__ data,
                                                             const float* __restrict__ g, float alpha, float eps, float lo,
                                                             float hi, float* delta_out, int64_t n) {
  __shared__ double s_scratch[32];
  __shared__ double s_part;
  const int64_t nvec = n / V, base = (int64_t)blockIdx.y * nvec;
  const Slice sl = my_slice(nvec);
  double acc = 0.0;
  for (int64_t i = sl.begin + threadIdx.x; i < sl.end; i += kThreads) {
    const Vec<V> v = ldv<V>(g, base + i);
#pragma unroll
    for (int k = 0; k < V; ++k) acc += (double)v.v[k] * (double)v.v[k];
  }
  const float gn = (float)sqrt(cluster_allreduce_sum(acc, s_scratch, &s_part));
  const float den = add_rn(gn, 1e-20f);
  acc = 0.0;
  for (int64_t i = sl.begin + threadIdx.x; i < sl.end; i += kThreads) {
    const Vec<V> gv = ldv<V>(g, base + i), dv = ldv_rw<V>(delta, base + i);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float y = add_rn(dv.v[k], mul_rn(div_rn(gv.v[k], den), alpha));
      acc += (double)y * (double)y;
    }
  }
  const float yn = (float)sqrt(cluster_allreduce_sum(acc, s_scratch, &s_part));
  const bool shrink = yn > eps;
  const float f = shrink ? div_rn(eps, add_rn(yn, 1e-7f)) : 1.0f;
  for (int64_t i = sl.begin + threadIdx.x; i < sl.end; i += kThreads) {
    const Vec<V> gv = ldv<V>(g, base + i), dv = ldv_rw<V>(delta, base + i), xv = ldv<V>(data, base + i);
    Vec<V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float y = add_rn(dv.v[k], mul_rn(div_rn(gv.v[k], den), alpha));
      if (shrink) y = mul_rn(y, f);
      o.v[k] = min_nan(max_nan(y, sub_rn(lo, xv.v[k])), sub_rn(hi, xv.v[k]));
    }
    stv<V>(delta_out, base + i, o);
  }
}

// attack.py:136-141
template <int V>
__global__ void __launch_bounds__(kThreads) init_l2_kernel(const float* delta, const float* __restrict__ r,
                                                           const float* __restrict__ data, float eps, float lo, float hi,
                                                           float* out, int64_t n) {
  __shared__ double s_scratch[32];
  __shared__ double s_part;
  const int64_t nvec = n / V, base = (int64_t)blockIdx.y * nvec;
  const Slice sl = my_slice(nvec);
  double acc = 0.0;
  for (int64_t i = sl.begin + threadIdx.x; i < sl.end; i += kThreads) {
    const Vec<V> v = ldv_rw<V>(delta, base + i);
#pragma unroll
    for (int k = 0; k < V; ++k) acc += (double)v.v[k] * (double)v.v[k];
  }
  const float nn = (float)sqrt(cluster_allreduce_sum(acc, s_scratch, &s_part));
  for (int64_t i = sl.begin + threadIdx.x; i < sl.end; i += kThreads) {
    const Vec<V> dv = ldv_rw<V>(delta, base + i), rv = ldv<V>(r, base + i), xv = ldv<V>(data, base + i);
    Vec<V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float f = mul_rn(div_rn(rv.v[k], nn), eps);
      o.v[k] = min_nan(max_nan(mul_rn(dv.v[k], f), sub_rn(lo, xv.v[k])), sub_rn(hi, xv.v[k]));
    }
    stv<V>(out, base + i, o);
  }
}

}  // namespace

extern "C" {

int64_t ta_abs_mean_ws_bytes(int B, int64_t n) {
  (void)n;
  return (B > 0) ? 0 : 0;   // cluster/DSMEM reduction needs no global scratch
}

int ta_abs_mean_per_sample(const float* g, float* mean_out, int B, int64_t n, int mode, void* ws, ta_stream_t stream) {
  (void)ws;
  TA_REQUIRE(g && mean_out && B > 0 && n > 0, "ta_abs_mean_per_sample: null pointer or empty shape (B=%d n=%lld)", B, (long long)n);
  if (mode == TA_MEAN_TORCH) return aten_abs_mean_launch(g, mean_out, B, n, nullptr, (cudaStream_t)stream);
  if (mode != TA_MEAN_EXACT) {
    set_error("ta_abs_mean_per_sample: mode %d not available in this build", mode);
    return TA_EUNSUPPORTED;
  }
  const bool v4 = (n % 4 == 0) && aligned16(g);
  const int cl = pick_cluster(n, kThreads);
  if (v4) return launch_cluster("ta_abs_mean_per_sample", abs_mean_kernel<4>, cl, B, kThreads, 0, (cudaStream_t)stream, g, mean_out, n);
  return launch_cluster("ta_abs_mean_per_sample", abs_mean_kernel<1>, cl, B, kThreads, 0, (cudaStream_t)stream, g, mean_out, n);
}

int64_t ta_update_l2_ws_bytes(int B) { (void)B; return 0; }

int ta_update_l2(const float* delta, const float* data, const float* g, float alpha, float eps, float lo, float hi,
                 float* delta_out, int B, int64_t n, void* ws, ta_stream_t stream) {
  (void)ws;
  TA_REQUIRE(delta && data && g && delta_out && B > 0 && n > 0, "ta_update_l2: null pointer or empty shape");
  const bool v4 = (n % 4 == 0) && aligned16(delta) && aligned16(data) && aligned16(g) && aligned16(delta_out);
  const int cl = pick_cluster(n, kThreads);
  if (v4) return launch_cluster("ta_update_l2", update_l2_kernel<4>, cl, B, kThreads, 0, (cudaStream_t)stream, delta, data, g, alpha, eps, lo, hi, delta_out, n);
  return launch_cluster("ta_update_l2", update_l2_kernel<1>, cl, B, kThreads, 0, (cudaStream_t)stream, delta, data, g, alpha, eps, lo, hi, delta_out, n);
}

int ta_init_l2_scale(const float* delta, const float* r, const float* data, float eps, float lo, float hi, float* out, int B,
                     int64_t n, void* ws, ta_stream_t stream) {
  (void)ws;
  TA_REQUIRE(delta && r && data && out && B > 0 && n > 0, "ta_init_l2_scale: null pointer or empty shape");
  const bool v4 = (n % 4 == 0) && aligned16(delta) && aligned16(r) && aligned16(data) && aligned16(out);
  const int cl = pick_cluster(n, kThreads);
  if (v4) return launch_cluster("ta_init_l2_scale", init_l2_kernel<4>, cl, B, kThreads, 0, (cudaStream_t)stream, delta, r, data, eps, lo, hi, out, n);
  return launch_cluster("ta_init_l2_scale", init_l2_kernel<1>, cl, B, kThreads, 0, (cudaStream_t)stream, delta, r, data, eps, lo, hi, out, n);
}

}  // extern "C"
