// What does an instruction of the trace kernels cost to ISSUE on a B200 SMSP?  (round 2)
// k_spot_rev did not speed up from 8 to 12 to 16 resident warps per SM, so the fused pass is bound
// by a per-instruction throughput limit, not by latency.  This probe times fixed instruction
// patterns (inline PTX, so the forms are pinned; check with cuobjdump -sass) with 1..8 warps per SMSP:
// cycles per instruction per SMSP for packed FFMA2 / FMUL2 / FADD2 with distinct, shared and
// scalar-broadcast operands, scalar FFMA, and mixes with MUFU / integer / scalar FP work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench4 tools/microbench4.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#define ITERS 2048
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float a, float b) {
  return ((u64)__float_as_uint(b) << 32) | (u64)__float_as_uint(a);
}
#define FFMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define FMUL2(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define FADD2(d, a, b) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define FFMA(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define MUFU(d) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(d))
#define IMAD(d, a, b) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b))

enum {
  P_FFMA2_DISTINCT,   // x[i] = x[i] * y[i] + z[i]
  P_FFMA2_ACC,        // x[i] = y[i] * z[i] + x[i]
  P_FFMA2_SHARED,     // x[i] = x[i] * y[0] + z[0]         (two operands shared by all eight)
  P_FFMA2_SQUARE,     // x[i] = y[i] * y[i] + x[i]         (a repeated operand)
  P_FMUL2,            // x[i] = x[i] * y[i]
  P_FADD2,            // x[i] = x[i] + y[i]
  P_FFMA_DISTINCT,    // 16 scalar: x = x * y + z
  P_FFMA_SHARED,      // 16 scalar: x = x * y0 + z0
  P_MIX_FFMA2_FFMA,   // 8 FFMA2 distinct + 8 scalar FFMA distinct
  P_MIX_FFMA2_IMAD,   // 8 FFMA2 distinct + 8 IMAD
  P_MIX_FFMA2_MUFU4,  // 8 FFMA2 distinct + 4 MUFU
  P_MIX_FFMA_MUFU4,   // 16 FFMA distinct + 4 MUFU
  P_FFMA2_HALF,       // 4 FFMA2 distinct + 8 scalar FFMA distinct (same flops as 8 FFMA2)
  P_COUNT
};
const char *kNames[P_COUNT] = {"8 FFMA2 x=x*y+z (distinct)", "8 FFMA2 x=y*z+x (accumulate)", "8 FFMA2 x=x*y0+z0 (shared)",
                               "8 FFMA2 x=y*y+x (square)", "8 FMUL2 x=x*y", "8 FADD2 x=x+y",
                               "16 FFMA x=x*y+z (distinct)", "16 FFMA x=x*y0+z0 (shared)",
                               "8 FFMA2 + 8 FFMA (distinct)", "8 FFMA2 + 8 IMAD", "8 FFMA2 + 4 MUFU",
                               "16 FFMA + 4 MUFU", "4 FFMA2 + 8 FFMA (distinct)"};
const int kInstr[P_COUNT] = {8, 8, 8, 8, 8, 8, 16, 16, 16, 16, 12, 20, 12};

template <int P>
__global__ void __launch_bounds__(128) k(float *out, long long *cycles, float seed) {
  u64 x[8], y[8], z[8];
  float sx[16], sy[16], sz[16];
  int ia[8], ib = threadIdx.x + 3;
  float m[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = pk(seed + i, seed - i);
    const float tid = 1e-9f * threadIdx.x;          // per-thread values: real registers, no immediates
    y[i] = pk(0.999f + 1e-5f * i + tid, 1.0001f - 1e-5f * i - tid);
    z[i] = pk(seed * 1e-3f * i + tid, seed * 1e-3f - tid);
    ia[i] = threadIdx.x * (i + 1);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    sx[i] = seed + i;
    sy[i] = 0.999f + 1e-5f * i + 1e-9f * threadIdx.x;
    sz[i] = seed * 1e-3f * i + 1e-9f * threadIdx.x;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = 1.5f + i + seed;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (P == P_FFMA2_DISTINCT || P == P_MIX_FFMA2_FFMA || P == P_MIX_FFMA2_IMAD || P == P_MIX_FFMA2_MUFU4)
        FFMA2(x[i], x[i], y[i], z[i]);
      if (P == P_FFMA2_HALF && i < 4) FFMA2(x[i], x[i], y[i], z[i]);
      if (P == P_FFMA2_ACC) FFMA2(x[i], y[i], z[i], x[i]);
      if (P == P_FFMA2_SHARED) FFMA2(x[i], x[i], y[0], z[0]);
      if (P == P_FFMA2_SQUARE) FFMA2(x[i], y[i], y[i], x[i]);
      if (P == P_FMUL2) FMUL2(x[i], x[i], y[i]);
      if (P == P_FADD2) FADD2(x[i], x[i], y[i]);
      if (P == P_MIX_FFMA2_FFMA || P == P_FFMA2_HALF) FFMA(sx[i], sx[i], sy[i], sz[i]);
      if (P == P_MIX_FFMA2_IMAD) IMAD(ia[i], ib, ia[(i + 1) & 7]);
      if ((P == P_MIX_FFMA2_MUFU4 || P == P_MIX_FFMA_MUFU4) && (i & 1)) MUFU(m[i >> 1]);
    }
    if (P == P_FFMA_DISTINCT || P == P_MIX_FFMA_MUFU4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) FFMA(sx[i], sx[i], sy[i], sz[i]);
    }
    if (P == P_FFMA_SHARED) {
#pragma unroll
      for (int i = 0; i < 16; ++i) FFMA(sx[i], sx[i], sy[0], sz[0]);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)x[i]) + __uint_as_float((unsigned)(x[i] >> 32)) + ia[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += sx[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) {      // per warp: SM id, start, stop (the clock is per SM)
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long *rec = cycles + 3 * (blockIdx.x * 4 + (threadIdx.x >> 5));
    rec[0] = smid;
    rec[1] = t0;
    rec[2] = t1;
  }
}

template <int P>
void run(int sms, float *out, long long *cyc_dev) {
  printf("%-34s", kNames[P]);
  for (int w : {1, 2, 4, 6}) {
    const int blocks = sms * w;      // 128-thread blocks: one warp per SMSP each -> w warps per SMSP
    k<P><<<blocks, 128>>>(out, cyc_dev, 0.5f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<P><<<blocks, 128>>>(out, cyc_dev, 0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    static long long host[3 * 148 * 8 * 4];
    cudaMemcpy(host, cyc_dev, sizeof(long long) * 3 * blocks * 4, cudaMemcpyDeviceToHost);
    // the scheduler is not fair between warps, so a warp's own elapsed time says little: take, per SM,
    // the span from the first start to the last stop of the warps that ran there, and how many ran there
    static long long lo[256], hi[256];
    static int count[256];
    for (int i = 0; i < 256; ++i) { lo[i] = 1LL << 62; hi[i] = 0; count[i] = 0; }
    for (int i = 0; i < blocks * 4; ++i) {
      const int sm = (int)host[3 * i];
      if (host[3 * i + 1] < lo[sm]) lo[sm] = host[3 * i + 1];
      if (host[3 * i + 2] > hi[sm]) hi[sm] = host[3 * i + 2];
      count[sm]++;
    }
    double cyc = 0, n = 0;
    int used = 0, max_count = 0;
    for (int i = 0; i < 256; ++i)
      if (count[i]) {
        cyc += (double)(hi[i] - lo[i]) / ((double)ITERS * kInstr[P] * count[i] / 4.0);   // per SMSP
        n += 1;
        used++;
        if (count[i] > max_count) max_count = count[i];
      }
    printf("  w=%d: %5.2f (%d SMs, <=%d warps/SM, %.0f us)", w, cyc / n, used, max_count, ms * 1e3);
  }
  printf("\n");
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  float *out;
  long long *cyc;
  cudaMalloc(&out, 1 << 24);
  cudaMalloc(&cyc, sizeof(long long) * 3 * sms * 8 * 4);
  printf("%s, %d SMs; cycles per instruction per SMSP (issue cost), ITERS=%d\n", p.name, sms, ITERS);
  run<P_FFMA2_DISTINCT>(sms, out, cyc);
  run<P_FFMA2_ACC>(sms, out, cyc);
  run<P_FFMA2_SHARED>(sms, out, cyc);
  run<P_FFMA2_SQUARE>(sms, out, cyc);
  run<P_FMUL2>(sms, out, cyc);
  run<P_FADD2>(sms, out, cyc);
  run<P_FFMA_DISTINCT>(sms, out, cyc);
  run<P_FFMA_SHARED>(sms, out, cyc);
  run<P_MIX_FFMA2_FFMA>(sms, out, cyc);
  run<P_MIX_FFMA2_IMAD>(sms, out, cyc);
  run<P_MIX_FFMA2_MUFU4>(sms, out, cyc);
  run<P_MIX_FFMA_MUFU4>(sms, out, cyc);
  run<P_FFMA2_HALF>(sms, out, cyc);
  return 0;
}
