// Microbenchmark: FP32 FMA issue rate per SM on sm_100a for (A) 3-register FFMA, (B) FFMA with a constant-bank /
// uniform-register multiplier, (C) packed FFMA2 (fma.rn.f32x2). Prints FMA lanes per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu && ./ffma_rate
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CHAINS = 16;

__global__ void k_reg(float* out, const float* in) {
  float a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  float acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = in[64 + i];
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(a), "f"(b));
  }
  float s = 0; 
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_const(float* out, const float* in, float ca, float cb) {
  float acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = in[64 + i];
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = fmaf(acc[i], ca, cb);      // multiplier / addend from the constant bank
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// sliding-window shape of the TIM row pass: acc[c] += w_j * v[j + c] with w uniform (param) and v per-thread registers
__global__ void k_window(float* out, const float* in, float w0, float w1, float w2, float w3) {
  float v[8], acc[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = in[(threadIdx.x & 31) + i];
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[c]) : "f"(w0), "f"(v[c]));
      asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[c]) : "f"(w1), "f"(v[c + 1]));
      asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[c]) : "f"(w2), "f"(v[c + 2]));
      asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[c]) : "f"(w3), "f"(v[c + 3]));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}
__global__ void k_f2(float* out, const float* in) {
  unsigned long long a, b, acc[CHAINS / 2];
  float a0 = in[threadIdx.x & 31], b0 = in[32 + (threadIdx.x & 31)];
  asm("mov.b64 %0, {%1,%1};" : "=l"(a) : "f"(a0));
  asm("mov.b64 %0, {%1,%1};" : "=l"(b) : "f"(b0));
#pragma unroll
  for (int i = 0; i < CHAINS / 2; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(acc[i]) : "f"(in[64 + 2 * i]), "f"(in[65 + 2 * i]));
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a), "l"(b));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS / 2; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i])); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float *in, *out; cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
  for (int wps : {4, 8, 16, 32}) {          // warps per SM
    const int threads = 128, blocks = sms * wps * 32 / threads;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    const double lanes = (double)blocks * threads * ITERS * CHAINS;
    float t;
    t = timeit([&] { k_reg<<<blocks, threads>>>(out, in); });
    printf("warps/SM %2d  FFMA 3-reg        : %7.1f FMA/clk/SM (at %d MHz nominal)  %.3f ms\n", wps, lanes / (t * 1e-3) / (khz * 1e3) / sms, khz / 1000, t);
    t = timeit([&] { k_const<<<blocks, threads>>>(out, in, 1.0f, 0.0f); });
    printf("warps/SM %2d  FFMA const/UR     : %7.1f FMA/clk/SM  %.3f ms\n", wps, lanes / (t * 1e-3) / (khz * 1e3) / sms, t);
    t = timeit([&] { k_window<<<blocks, threads>>>(out, in, 1.0f, 0.5f, 0.25f, 0.125f); });
    printf("warps/SM %2d  FFMA window (UR w): %7.1f FMA/clk/SM  %.3f ms\n", wps, lanes / (t * 1e-3) / (khz * 1e3) / sms, t);
    t = timeit([&] { k_f2<<<blocks, threads>>>(out, in); });
    printf("warps/SM %2d  FFMA2 (f32x2)     : %7.1f FMA/clk/SM  %.3f ms\n", wps, lanes / (t * 1e-3) / (khz * 1e3) / sms, t);
    cudaFree(out);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
