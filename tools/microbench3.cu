// Do non-FMA instructions issue in the shadow of packed FFMA2 when its operands are three
// distinct register pairs (no operand reuse)?  8 FFMA2 + N integer ops per iteration.
#include <cuda_runtime.h>
#include <stdio.h>
#define ITERS 8192

template <int NALU, bool PACKED>
__global__ void k(float *out, float a, int salt) {
  float2 x[8], y[8], z[8];
  int acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = make_float2(threadIdx.x + i, 1.f + i);
    y[i] = make_float2(0.999f + 1e-4f * i, 1.0001f);
    z[i] = make_float2(a * i, a);
    acc[i] = threadIdx.x * (i + 1);
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) {
        x[i] = __ffma2_rn(x[i], y[(i + 1) & 7], z[(i + 2) & 7]);
      } else {
        x[i].x = __fmaf_rn(x[i].x, y[(i + 1) & 7].x, z[(i + 2) & 7].x);
        x[i].y = __fmaf_rn(x[i].y, y[(i + 1) & 7].y, z[(i + 2) & 7].y);
      }
      if (i < NALU) acc[i] = (acc[i] ^ salt) + (acc[(i + 1) & 7] & it);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NALU, bool PACKED>
void run(const char *name, int sms, int wps, float *out, double ghz) {
  const int blocks = sms * wps;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NALU, PACKED><<<blocks, 128>>>(out, 0.5f, 3);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 3; ++i) k<NALU, PACKED><<<blocks, 128>>>(out, 0.5f, 3);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  printf("%-40s warps/SMSP=%d  %.3f ms  %.1f cycles per iteration per warp-slot\n", name, wps, ms,
         ms * 1e-3 * ghz * 1e9 / ITERS / wps);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float *out; cudaMalloc(&out, 1 << 24);
  const double ghz = p.clockRate * 1e-6;
  const int sms = p.multiProcessorCount;
  for (int w : {2, 4, 8}) {
    run<0, true>("8 FFMA2 (distinct operands)", sms, w, out, ghz);
    run<4, true>("8 FFMA2 + 4x(LOP3,LOP3,IADD)", sms, w, out, ghz);
    run<8, true>("8 FFMA2 + 8x(LOP3,LOP3,IADD)", sms, w, out, ghz);
    run<0, false>("16 FFMA scalar", sms, w, out, ghz);
    run<8, false>("16 FFMA scalar + 8x(LOP3,LOP3,IADD)", sms, w, out, ghz);
  }
  return 0;
}
