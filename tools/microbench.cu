// Issue-rate probes for the design of the trace kernels (B200, sm_100a):
//   scalar FFMA vs packed FFMA2 (fma.rn.f32x2), and how many non-FMA instructions
//   (MUFU, FSEL/compare, shared-memory loads) ride along for free.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 4096

__global__ void k_ffma(float *out, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __fmaf_rn(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float *out, float a, float b) {
  float2 x[8];
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8 FFMA2 + NMUFU rsqrt + NSEL selects per iteration
template <int NMUFU, int NSEL, int NLDS>
__global__ void k_mix2(float *out, float a, float b) {
  __shared__ float sh[1024];
  sh[threadIdx.x] = a;
  __syncthreads();
  float2 x[8];
  float m[4] = {1.5f + threadIdx.x, 2.5f, 3.5f, 4.5f};
  float sel[4] = {0.f, 1.f, 2.f, 3.f};
  float lds = 0.f;
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
#pragma unroll
    for (int i = 0; i < NMUFU; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
#pragma unroll
    for (int i = 0; i < NSEL; ++i) sel[i] = (x[i].x > b) ? sel[i] : x[i].y;
#pragma unroll
    for (int i = 0; i < NLDS; ++i) lds += sh[(threadIdx.x + it + i * 32) & 1023];
  }
  float s = lds;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += m[i] + sel[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 4, threads = 512;
  float *out;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  const double ghz = prop.clockRate * 1e-6;   // max SM clock
  const double lanes = (double)blocks * threads;
  auto report = [&](const char *name, float ms, double fma_per_thread_iter) {
    const double fma = lanes * ITERS * fma_per_thread_iter;
    printf("%-34s %8.3f ms  %7.2f TFLOP/s  %6.1f FMA-lanes/clk/SM (at %.3f GHz)\n", name, ms,
           2 * fma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / (ghz * 1e9) / sms, ghz);
  };
  report("FFMA  x8", time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 8);
  report("FFMA2 x8", time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16);
  report("FFMA2 x8 + 1 MUFU", time_ms([&] { k_mix2<1, 0, 0><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16);
  report("FFMA2 x8 + 2 MUFU", time_ms([&] { k_mix2<2, 0, 0><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16);
  report("FFMA2 x8 + 4 MUFU", time_ms([&] { k_mix2<4, 0, 0><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16);
  report("FFMA2 x8 + 4 SEL", time_ms([&] { k_mix2<0, 4, 0><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16);
  report("FFMA2 x8 + 4 LDS", time_ms([&] { k_mix2<0, 0, 4><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16);
  report("FFMA2 x8 + 2 MUFU + 4 SEL + 2 LDS", time_ms([&] { k_mix2<2, 4, 2><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16);
  cudaFree(out);
  return 0;
}
