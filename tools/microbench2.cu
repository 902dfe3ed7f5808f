// Packed-math operand-bandwidth probes (B200): does FFMA2 / FMUL2 / FADD2 keep its
// 2-cycle rate when all three operands are distinct register pairs (no .reuse)?
// How do scalar FFMA with distinct operands compare?  Warps per SMSP is a parameter.
#include <cuda_runtime.h>
#include <stdio.h>
#define ITERS 2048

template <int KIND>
__global__ void k(float *out, float a) {
  float2 x[6], y[6], z[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    x[i] = make_float2(threadIdx.x + i, 1.f + i);
    y[i] = make_float2(0.999f + 1e-4f * i, 1.0001f);
    z[i] = make_float2(a * i, a);
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (KIND == 0) x[i] = __ffma2_rn(x[i], y[(i + 1) % 6], z[(i + 2) % 6]);      // 3 distinct pairs
      if (KIND == 1) x[i] = __fmul2_rn(x[i], y[(i + 1) % 6]);
      if (KIND == 2) x[i] = __fadd2_rn(x[i], z[(i + 1) % 6]);
      if (KIND == 3) { x[i].x = __fmaf_rn(x[i].x, y[(i + 1) % 6].x, z[(i + 2) % 6].x);
                       x[i].y = __fmaf_rn(x[i].y, y[(i + 1) % 6].y, z[(i + 2) % 6].y); }
      if (KIND == 4) x[i] = __ffma2_rn(x[i], y[(i + 1) % 6], x[(i + 3) % 6]);      // chained across regs
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char *name, int sms, int warps_per_smsp, float *out, double ghz) {
  const int threads = 128, blocks = sms * warps_per_smsp;   // 4 warps per block, one per SMSP
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<blocks, threads>>>(out, 0.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) k<KIND><<<blocks, threads>>>(out, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double inst = (double)ITERS * 6 * (KIND == 3 ? 2 : 1);       // warp-instr per warp
  const double cyc = ms * 1e-3 * ghz * 1e9;
  printf("%-28s warps/SMSP=%2d  %.3f ms  %.2f cycles per warp-instruction per SMSP\n", name,
         warps_per_smsp, ms, cyc / (inst * warps_per_smsp));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float *out; cudaMalloc(&out, 1 << 24);
  const double ghz = p.clockRate * 1e-6;
  for (int w : {1, 2, 4, 8}) {
    run<0>("FFMA2 3 distinct pairs", p.multiProcessorCount, w, out, ghz);
    run<4>("FFMA2 chained regs", p.multiProcessorCount, w, out, ghz);
    run<1>("FMUL2 2 distinct pairs", p.multiProcessorCount, w, out, ghz);
    run<2>("FADD2 2 distinct pairs", p.multiProcessorCount, w, out, ghz);
    run<3>("FFMA scalar distinct", p.multiProcessorCount, w, out, ghz);
  }
  return 0;
}
