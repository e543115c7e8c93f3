// philox.cu — VMI-FGSM's neighbour staging with the noise generated in the kernel (gradient/vmifgsm.py:50,56):
//   reference per neighbour:  noise = zeros_like(delta).uniform_(-r, r)   (2 launches, 8 B/elem written)
//                             x_near = data + delta + noise                (2 launches, 24 B/elem)
//   here:                     out = (data + delta) + noise(i) [+ coef * look]   one launch, 12 B/elem (16 with look),
// where noise(i) is EXACTLY the value torch's CUDA uniform_ would have written at element i for the generator state
// (seed, offset) the caller read from torch's device generator — so the attack consumes the same random stream and stays
// bit-identical to the reference — followed by the caller advancing the generator by ta_uniform_fill_policy's increment.
//
// torch (ATen/native/cuda/DistributionTemplates.h, 2.11) fills a contiguous tensor with T = 256 * grid threads,
// grid = min(#SM * (maxThreadsPerSM / 256), ceil(numel / 256)); thread idx initialises Philox4_32_10 with
// (seed, subsequence = idx, offset) and its j-th curand_uniform4 call supplies elements idx + T * (4j + ii), ii = 0..3:
//   counter = (offset / 4 + j  [low 64 bits], idx [high 64 bits]), key = seed;  u = float(x_ii) * 2^-32 + 2^-33;
//   value = fma(u, to - from, from);  value == to -> from.
// One thread of this kernel owns 4 consecutive idx and one j: 4 Philox evaluations (≈ 15 integer instructions per element,
// far under the HBM time of 12 B) and four 128-bit load / store groups at stride T.
#include "common.cuh"

using namespace ta;

namespace {

constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kM0, c.x), lo0 = kM0 * c.x;
    const uint32_t hi1 = __umulhi(kM1, c.z), lo1 = kM1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    if (r < 9) { k.x += kW0; k.y += kW1; }
  }
  return c;
}

struct PhiloxStage {
  const float* data; const float* delta; const float* look; float* out; float* noise_out;
  float coef, from, to, range;
  uint64_t seed, ctr0;        // ctr0 = offset / 4
  int64_t N, T;
  int fma;
};

__device__ __forceinline__ float uniform_value(uint32_t x, const PhiloxStage& p) {
  const float u = add_rn(mul_rn((float)x, 2.3283064365386963e-10f), 1.1641532182693481e-10f);   // x * 2^-32 + 2^-33
  const float v = p.fma ? fmaf(u, p.range, p.from) : add_rn(mul_rn(u, p.range), p.from);
  return v == p.to ? p.from : v;
}

__device__ __forceinline__ float stage_value(float x, float d, float n, float l, bool has_look, float coef) {
  float r = add_rn(add_rn(x, d), n);
  if (has_look) r = add_rn(r, mul_rn(coef, l));
  return r;
}

// grid.x covers T / 4 groups of 4 consecutive thread indices, grid.y = j (Philox call number)
template <bool VEC>
__global__ void __launch_bounds__(256) neighbor_stage_philox_kernel(const PhiloxStage p) {
  const int64_t g4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // group of 4 torch-thread indices
  const int64_t idx0 = g4 * 4;
  if (idx0 >= p.T) return;
  const uint64_t j = blockIdx.y;
  const uint64_t ctr = p.ctr0 + j;
  const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
  uint4 r[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const uint64_t idx = (uint64_t)(idx0 + t);
    r[t] = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)idx, (uint32_t)(idx >> 32)), key);
  }
  const bool has_look = p.look != nullptr;
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const int64_t li = idx0 + p.T * (int64_t)(4 * j + ii);               // element of lane t = 0; lanes are consecutive
    if (li >= p.N) continue;
    const uint32_t xs[4] = {ii == 0 ? r[0].x : ii == 1 ? r[0].y : ii == 2 ? r[0].z : r[0].w,
                            ii == 0 ? r[1].x : ii == 1 ? r[1].y : ii == 2 ? r[1].z : r[1].w,
                            ii == 0 ? r[2].x : ii == 1 ? r[2].y : ii == 2 ? r[2].z : r[2].w,
                            ii == 0 ? r[3].x : ii == 1 ? r[3].y : ii == 2 ? r[3].z : r[3].w};
    if (VEC && idx0 + 3 < p.T && li + 3 < p.N) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(p.data + li));
      const float4 d = __ldg(reinterpret_cast<const float4*>(p.delta + li));
      float4 l = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_look) l = __ldg(reinterpret_cast<const float4*>(p.look + li));
      const float4 n = make_float4(uniform_value(xs[0], p), uniform_value(xs[1], p), uniform_value(xs[2], p), uniform_value(xs[3], p));
      const float4 o = make_float4(stage_value(x.x, d.x, n.x, l.x, has_look, p.coef), stage_value(x.y, d.y, n.y, l.y, has_look, p.coef),
                                   stage_value(x.z, d.z, n.z, l.z, has_look, p.coef), stage_value(x.w, d.w, n.w, l.w, has_look, p.coef));
      *reinterpret_cast<float4*>(p.out + li) = o;
      if (p.noise_out) *reinterpret_cast<float4*>(p.noise_out + li) = n;
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int64_t e = li + t;
        if (idx0 + t < p.T && e < p.N) {
          const float n = uniform_value(xs[t], p);
          p.out[e] = stage_value(__ldg(p.data + e), __ldg(p.delta + e), n, has_look ? __ldg(p.look + e) : 0.0f, has_look, p.coef);
          if (p.noise_out) p.noise_out[e] = n;
        }
      }
    }
  }
}

int torch_policy(int64_t numel, int64_t* T, int64_t* incr) {
  int dev = 0, sms = 0, mt = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&mt, cudaDevAttrMaxThreadsPerMultiProcessor, dev) != cudaSuccess) {
    set_error("ta_uniform_fill_policy: cannot query the device");
    cudaGetLastError();
    return TA_ECUDA;
  }
  const int64_t block = 256;
  int64_t grid = (numel + block - 1) / block;
  const int64_t cap = (int64_t)sms * (mt / block);
  if (grid > cap) grid = cap;
  *T = block * grid;
  *incr = ((numel - 1) / (*T * 4) + 1) * 4;
  return TA_OK;
}

}  // namespace

extern "C" {

int ta_uniform_fill_policy(int64_t numel, int64_t* threads_total, int64_t* offset_increment) {
  TA_REQUIRE(numel > 0 && threads_total && offset_increment, "ta_uniform_fill_policy: bad arguments");
  return torch_policy(numel, threads_total, offset_increment);
}

int ta_neighbor_stage_philox(const float* data, const float* delta, const float* look, float coef, float from, float to,
                             uint64_t seed, uint64_t offset, float* out, float* noise_out, int64_t N, ta_stream_t stream) {
  TA_REQUIRE(data && delta && out && N > 0, "ta_neighbor_stage_philox: null pointer or N=%lld", (long long)N);
  TA_REQUIRE(offset % 4 == 0, "ta_neighbor_stage_philox: philox offset %llu is not a multiple of 4", (unsigned long long)offset);
  TA_REQUIRE(N <= 0x7fffffffLL, "ta_neighbor_stage_philox: N=%lld needs torch's split 32-bit indexing (not supported)", (long long)N);
  int64_t T = 0, incr = 0;
  const int rc = torch_policy(N, &T, &incr);
  if (rc != TA_OK) return rc;
  PhiloxStage p{data, delta, look, out, noise_out, coef, from, to, (float)(to - from), seed, offset / 4, N, T,
                tune_get("philox.fma", 1)};
  const int64_t groups = (T + 3) / 4;
  const int64_t calls = incr / 4;                                   // Philox calls per torch thread
  TA_REQUIRE(calls <= 65535, "ta_neighbor_stage_philox: %lld draws per thread exceed the grid", (long long)calls);
  const bool vec = (T % 4 == 0) && aligned16(data) && aligned16(delta) && aligned16(out) && aligned16(look) && aligned16(noise_out);
  dim3 grid((unsigned)((groups + 255) / 256), (unsigned)calls);
  if (vec) neighbor_stage_philox_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else neighbor_stage_philox_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  return check_launch("ta_neighbor_stage_philox");
}

}  // extern "C"
