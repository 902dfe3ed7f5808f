// sm_100a kernels of the sequential ray-trace hot path and their C ABI
// (include/torchoptics_b200.h).  The per-ray arithmetic lives in trace_core.cuh;
// this file maps rays to threads, parks per-surface ray state in shared memory for
// the adjoint sweep, keeps per-thread gradient accumulators in registers and
// reduces them (warp shuffles -> shared memory -> fp64 partials -> tiny
// deterministic finalize kernels).
//
// Work decomposition used by every trace kernel: one CTA works on one (lens b,
// field f, wavelength w) and one contiguous chunk of the pupil axis, so that the
// surface table of that (b, w) sits in shared memory and every gradient
// accumulator of a thread belongs to one (b, f, w).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/torchoptics_b200.h"
#include "trace_core.cuh"

using namespace tl;

namespace {

thread_local char g_error[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char *fmt, const char *detail = "") {
  snprintf(g_error, sizeof(g_error), fmt, detail);
  return code;
}

#define TL_CHECK_CUDA(expr)                                                      \
  do {                                                                           \
    cudaError_t err__ = (expr);                                                  \
    if (err__ != cudaSuccess) return fail(TL_ERR_CUDA, #expr ": %s", cudaGetErrorString(err__)); \
  } while (0)

constexpr int kTraceThreads = 128;   // adjoint kernels: 6*S floats of state per thread in smem
constexpr int kFwdThreads = 256;
constexpr int kMaxRaysPerThread = 128;  // bounds the fp32 run length of an accumulator

struct DeviceInfo {
  int device = -1;
  int sms = 0;
};

int device_info(DeviceInfo &info) {
  int dev = 0;
  TL_CHECK_CUDA(cudaGetDevice(&dev));
  static thread_local DeviceInfo cached;
  if (cached.device != dev) {
    cached.device = dev;
    TL_CHECK_CUDA(cudaDeviceGetAttribute(&cached.sms, cudaDevAttrMultiProcessorCount, dev));
  }
  info = cached;
  return TL_OK;
}

// --------------------------------------------------------------------------
// device helpers
// --------------------------------------------------------------------------
struct Coord {
  int b, f, w;
};

__device__ __forceinline__ int64_t offset_of(const TlStrided &s, int b, int f, int p, int w) {
  return (int64_t)b * s.stride[0] + (int64_t)f * s.stride[1] + (int64_t)p * s.stride[2] +
         (int64_t)w * s.stride[3];
}

// Surface table of one (lens, wavelength) in shared memory.
struct Table {
  float *c, *t, *mu, *mu2;
  int *live;
  float length;   // sum |t|
};

__device__ __forceinline__ size_t table_floats(int S) { return 5 * (size_t)S; }

__device__ __forceinline__ Table load_table(float *base, const TlProblem &pb, int b, int w) {
  Table tab;
  const int S = pb.S;
  tab.c = base;
  tab.t = base + S;
  tab.mu = base + 2 * S;
  tab.mu2 = base + 3 * S;
  tab.live = reinterpret_cast<int *>(base + 4 * S);
  for (int k = threadIdx.x; k < S; k += blockDim.x) {
    const float m = pb.mu[((int64_t)b * pb.W + w) * S + k];
    tab.c[k] = pb.c[(int64_t)b * S + k];
    tab.t[k] = pb.t[(int64_t)b * S + k];
    tab.mu[k] = m;
    tab.mu2[k] = m * m;
    tab.live[k] = pb.live[(int64_t)b * S + k] != 0;
  }
  __syncthreads();
  float len = 0.f;
  for (int k = 0; k < S; ++k) len += fabsf(tab.t[k]);
  tab.length = len;
  return tab;
}

struct Traced {
  Ray<float> pre;    // state in front of the image plane (after the last z shift)
  float x, y;        // image-plane point
  bool ok, backward;
};

// Per-thread slot of the parked in-states: element j of surface k lives at
// state[(k * 6 + j) * stride] (stride = threads per CTA -> conflict-free).
template <bool SAVE>
__device__ __forceinline__ void park_state(float *state, int stride, int k, const Ray<float> &r) {
  if (SAVE) {
    float *s = state + (size_t)k * 6 * stride;
    s[0] = r.x;
    s[stride] = r.y;
    s[2 * stride] = r.z;
    s[3 * stride] = r.cx;
    s[4 * stride] = r.cy;
    s[5 * stride] = r.cz;
  }
}

__device__ __forceinline__ Ray<float> load_state(const float *state, int stride, int k) {
  const float *s = state + (size_t)k * 6 * stride;
  Ray<float> r;
  r.x = s[0];
  r.y = s[stride];
  r.z = s[2 * stride];
  r.cx = s[3 * stride];
  r.cy = s[4 * stride];
  r.cz = s[5 * stride];
  return r;
}

// Exact-policy trace of one ray (rtl:594-675 statement by statement).
template <bool SAVE>
__device__ __noinline__ Traced trace_exact(float x, float y, float z, float cx, float cy,
                                           const Table &tab, int S, bool allow_backward,
                                           float *state, int stride) {
  Ray<float> r{x, y, z, cx, cy, exact_cz0(cx, cy)};
  bool ok = true, backward = false;
  for (int k = 0; k < S; ++k) {
    park_state<SAVE>(state, stride, k, r);
    const Surface s{tab.c[k], tab.t[k], tab.mu[k]};
    exact_surface(r, s, k > 0 && tab.live[k - 1], allow_backward, ok, backward);
  }
  Traced out;
  out.pre = r;
  exact_image(r, tab.live[S - 1] != 0, allow_backward, ok, backward);
  out.x = r.x;
  out.y = r.y;
  out.ok = ok;
  out.backward = backward;
  return out;
}

// Guarded policy: contracted fast path, exact re-trace unless clearly good.
template <bool SAVE>
__device__ __forceinline__ Traced trace_guarded(float x, float y, float z, float cx, float cy,
                                                const Table &tab, int S, bool allow_backward,
                                                int arith, float *state, int stride) {
  if (arith == TL_ARITH_GUARDED) {
    Ray<float> r{x, y, z, cx, cy, fast_cz0(cx, cy)};
    float min_cos2 = 1.0f, min_travel = 3.0e38f;
    for (int k = 0; k < S; ++k) {
      park_state<SAVE>(state, stride, k, r);
      float travel;
      fast_surface(r, tab.c[k], tab.mu[k], tab.mu2[k], tab.t[k], min_cos2, travel);
      if (k > 0 && tab.live[k - 1]) min_travel = fminf(min_travel, travel);
    }
    Traced out;
    out.pre = r;
    const float travel = fast_image(r);
    if (tab.live[S - 1]) min_travel = fminf(min_travel, travel);
    const float band = kBandTravelRel * fmaxf(1.0f, tab.length + fabsf(z));
    const float probe = r.x + r.y + r.cx + r.cy;
    const bool clear = (min_cos2 > kGuard + kBandCos2) && (min_travel > band) &&
                       (fabsf(probe) < 3.0e38f);
    if (clear) {
      out.x = r.x;
      out.y = r.y;
      out.ok = true;
      out.backward = false;
      return out;
    }
  }
  return trace_exact<SAVE>(x, y, z, cx, cy, tab, S, allow_backward, state, stride);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  return v;
}

// --------------------------------------------------------------------------
// K1: forward trace (trace_skew, rtl:594-675)
// grid: (b, f, w, chunk) flattened; CTA strides over its pupil chunk.
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(kFwdThreads)
k_trace_fwd(TlProblem pb, TlTraceOut out, int nchunks, int chunk_len) {
  extern __shared__ float smem[];
  int blk = blockIdx.x;
  const int chunk = blk % nchunks; blk /= nchunks;
  const int w = blk % pb.W; blk /= pb.W;
  const int f = blk % pb.F;
  const int b = blk / pb.F;
  const Table tab = load_table(smem, pb, b, w);
  const int p_lo = chunk * chunk_len;
  const int p_hi = min(pb.P, p_lo + chunk_len);
  for (int p = p_lo + threadIdx.x; p < p_hi; p += blockDim.x) {
    const float x = pb.x.ptr[offset_of(pb.x, b, f, p, w)];
    const float y = pb.y.ptr[offset_of(pb.y, b, f, p, w)];
    const float z = pb.z.ptr[offset_of(pb.z, b, f, p, w)];
    const float cx = pb.cx.ptr[offset_of(pb.cx, b, f, p, w)];
    const float cy = pb.cy.ptr[offset_of(pb.cy, b, f, p, w)];
    const Traced tr = trace_guarded<false>(x, y, z, cx, cy, tab, pb.S,
                                           pb.allow_backward_rays != 0, pb.arith, nullptr, 0);
    const int64_t o = (((int64_t)b * pb.F + f) * pb.P + p) * pb.W + w;
    out.x[o] = tr.x;
    out.y[o] = tr.y;
    out.cx[o] = tr.pre.cx;
    out.cy[o] = tr.pre.cy;
    out.ok[o] = tr.ok;
    out.backward[o] = tr.backward;
  }
}

// --------------------------------------------------------------------------
// K2/K3: forward + adjoint in one pass.
//   MODE_BWD        seeds come from the caller (autograd of trace_skew)
//   MODE_SPOT_GRAD  unit seed on y; accumulates both sum(J) and sum((y-y0) J) so
//                   that the RMS gradient is assembled after the reduction
//   MODE_SPOT_EVAL  forward moments only (no state, no adjoint)
// --------------------------------------------------------------------------
enum { MODE_BWD = 0, MODE_SPOT_GRAD = 1, MODE_SPOT_EVAL = 2 };

struct AdjArgs {
  TlSeeds seeds;
  TlGrads grads;
  double *partial;   // [n_blocks, n_acc]
  float *ref_y;      // [B,F] (spot modes)
  int nchunks, chunk_len, n_acc;
};

template <int NS_MAX, int MODE>
__global__ void __launch_bounds__(kTraceThreads)
k_trace_adj(TlProblem pb, AdjArgs args) {
  extern __shared__ float smem[];
  constexpr bool kSpot = MODE != MODE_BWD;
  constexpr bool kAdjoint = MODE != MODE_SPOT_EVAL;
  constexpr int NA = kAdjoint ? NS_MAX : 1;
  const int S = pb.S;
  const int tid = threadIdx.x;
  int blk = blockIdx.x;
  const int chunk = blk % args.nchunks; blk /= args.nchunks;
  const int w = blk % pb.W; blk /= pb.W;
  const int f = blk % pb.F;
  const int b = blk / pb.F;
  const Table tab = load_table(smem, pb, b, w);
  float *state = smem + table_floats(S) + tid;
  const int stride = kTraceThreads;
  const bool allow_backward = pb.allow_backward_rays != 0;

  // reference height of this field: the exact-policy chief ray (pupil centre) of
  // wavelength 0 -- every CTA and every rank computes the same value.
  float y0 = 0.f;
  if (kSpot) {
    __shared__ float s_y0;
    if (tid == 0) {
      Ray<float> r{0.f, 0.f, pb.z.ptr[offset_of(pb.z, b, f, 0, 0)],
                   pb.cx.ptr[offset_of(pb.cx, b, f, 0, 0)],
                   pb.cy.ptr[offset_of(pb.cy, b, f, 0, 0)], 0.f};
      r.cz = exact_cz0(r.cx, r.cy);
      bool ok = true, backward = false;
      for (int k = 0; k < S; ++k) {
        const Surface s{tab.c[k], tab.t[k], pb.mu[((int64_t)b * pb.W) * S + k]};   // wavelength 0
        exact_surface(r, s, k > 0 && tab.live[k - 1], allow_backward, ok, backward);
      }
      exact_image(r, tab.live[S - 1] != 0, allow_backward, ok, backward);
      const float v = (ok && fabsf(r.y) < 3.0e38f) ? r.y : 0.f;
      s_y0 = v;
      if (chunk == 0 && w == 0) args.ref_y[b * pb.F + f] = v;
    }
    __syncthreads();
    y0 = s_y0;
  }

  // accumulators (registers: every index below is a compile-time constant)
  float acc_c[NA], acc_t[NA], acc_mu[NA];      // sum J        (BWD: sum of gradients)
  float wac_c[NA], wac_t[NA], wac_mu[NA];      // sum (y-y0) J (SPOT_GRAD only)
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    acc_c[k] = acc_t[k] = acc_mu[k] = 0.f;
    wac_c[k] = wac_t[k] = wac_mu[k] = 0.f;
  }
  float acc_z = 0.f, wac_z = 0.f;
  float m_s1 = 0.f, m_s2 = 0.f, m_n = 0.f;

  const int p_lo = pb.p_begin + chunk * args.chunk_len;
  const int p_hi = min(pb.p_end, p_lo + args.chunk_len);
  for (int p = p_lo + tid; p < p_hi; p += kTraceThreads) {
    const float x = pb.x.ptr[offset_of(pb.x, b, f, p, w)];
    const float y = pb.y.ptr[offset_of(pb.y, b, f, p, w)];
    const float z = pb.z.ptr[offset_of(pb.z, b, f, p, w)];
    const float cx = pb.cx.ptr[offset_of(pb.cx, b, f, p, w)];
    const float cy = pb.cy.ptr[offset_of(pb.cy, b, f, p, w)];
    const Traced tr = trace_guarded<kAdjoint>(x, y, z, cx, cy, tab, S, allow_backward, pb.arith,
                                              state, stride);
    const int64_t o = (((int64_t)b * pb.F + f) * pb.P + p) * pb.W + w;
    float wgt = 0.f;
    if (kSpot && tr.ok) {
      wgt = tr.y - y0;
      m_s1 += wgt;
      m_s2 = ffma(wgt, wgt, m_s2);
      m_n += 1.0f;
    }
    if (kAdjoint) {
      Ray<float> a{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (tr.ok) {
        float sx = 0.f, sy = 0.f, scx = 0.f, scy = 0.f;
        if (MODE == MODE_SPOT_GRAD) {
          sy = 1.0f;
        } else {
          if (args.seeds.gx) sx = args.seeds.gx[o];
          if (args.seeds.gy) sy = args.seeds.gy[o];
          if (args.seeds.gcx) scx = args.seeds.gcx[o];
          if (args.seeds.gcy) scy = args.seeds.gcy[o];
        }
        a = adjoint_image(tr.pre, sx, sy, scx, scy);
        Ray<float> next = tr.pre;
#pragma unroll
        for (int k = NS_MAX - 1; k >= 0; --k) {
          if (k < S) {
            const Ray<float> in = load_state(state, stride, k);
            const SurfaceGrad<float> g =
                adjoint_surface(in, next, tab.c[k], tab.mu[k], tab.mu2[k], a);
            acc_c[k] += g.c;
            acc_t[k] += g.t;
            acc_mu[k] += g.mu;
            if (MODE == MODE_SPOT_GRAD) {
              wac_c[k] = ffma(wgt, g.c, wac_c[k]);
              wac_t[k] = ffma(wgt, g.t, wac_t[k]);
              wac_mu[k] = ffma(wgt, g.mu, wac_mu[k]);
            }
            next = in;
          }
        }
        adjoint_cz0(next, a);
        acc_z += a.z;
        if (MODE == MODE_SPOT_GRAD) wac_z = ffma(wgt, a.z, wac_z);
      }
      if (MODE == MODE_BWD) {
        if (args.grads.gx) args.grads.gx[o] = a.x;
        if (args.grads.gy) args.grads.gy[o] = a.y;
        if (args.grads.gz) args.grads.gz[o] = a.z;
        if (args.grads.gcx) args.grads.gcx[o] = a.cx;
        if (args.grads.gcy) args.grads.gcy[o] = a.cy;
      }
    }
  }

  // ---- CTA reduction: warp shuffles -> smem [warp][slot] -> fp64 partial row
  __syncthreads();
  float *red = smem + table_floats(S);
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kTraceThreads / 32;
  const int n_acc = args.n_acc;
  auto put = [&](int slot, float v) {
    v = warp_sum(v);
    if (lane == 0) red[warp * n_acc + slot] = v;
  };
  if (MODE == MODE_BWD) {
#pragma unroll
    for (int k = 0; k < NA; ++k)
      if (k < S) {
        put(k, acc_c[k]);
        put(S + k, acc_t[k]);
        put(2 * S + k, acc_mu[k]);
      }
    put(3 * S, acc_z);
  } else if (MODE == MODE_SPOT_GRAD) {
#pragma unroll
    for (int k = 0; k < NA; ++k)
      if (k < S) {
        put(k, wac_c[k]);
        put(S + k, acc_c[k]);
        put(2 * S + k, wac_t[k]);
        put(3 * S + k, acc_t[k]);
        put(4 * S + k, wac_mu[k]);
        put(5 * S + k, acc_mu[k]);
      }
    put(6 * S, wac_z);
    put(6 * S + 1, acc_z);
    put(6 * S + 2, m_s1);
    put(6 * S + 3, m_s2);
    put(6 * S + 4, m_n);
  } else {
    put(0, m_s1);
    put(1, m_s2);
    put(2, m_n);
  }
  __syncthreads();
  for (int i = tid; i < n_acc; i += kTraceThreads) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) s += (double)red[q * n_acc + i];
    args.partial[(int64_t)blockIdx.x * n_acc + i] = s;
  }
}

// partial[(bfw * nchunks + chunk), n_acc] -> dst[bfw, n_acc], fixed summation order
__global__ void k_reduce_chunks(const double *partial, double *dst, int n_rows, int nchunks,
                                int n_acc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_rows * n_acc) return;
  const int row = (int)(i / n_acc), slot = (int)(i % n_acc);
  double s = 0.0;
  for (int c = 0; c < nchunks; ++c) s += partial[((int64_t)row * nchunks + c) * n_acc + slot];
  dst[i] = s;
}

// rows[b,f,w][3S+1] -> gc[b,S], gt[b,S], gmu[b,w,S], gz[b]
__global__ void k_bwd_finalize(const double *rows, TlGrads g, int B, int F, int W, int S) {
  const int n_acc = 3 * S + 1;
  const int per_lens = 2 * S + W * S + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * per_lens) return;
  const int b = i / per_lens;
  int j = i % per_lens;
  double s = 0.0;
  if (j < 2 * S) {                       // c (j < S) or t
    for (int f = 0; f < F; ++f)
      for (int w = 0; w < W; ++w) s += rows[(((int64_t)b * F + f) * W + w) * n_acc + j];
    if (j < S) g.gc[b * S + j] = (float)s;
    else g.gt[b * S + (j - S)] = (float)s;
  } else if (j < 2 * S + W * S) {
    j -= 2 * S;
    const int w = j / S, k = j % S;
    for (int f = 0; f < F; ++f) s += rows[(((int64_t)b * F + f) * W + w) * n_acc + 2 * S + k];
    g.gmu[((int64_t)b * W + w) * S + k] = (float)s;
  } else {
    for (int f = 0; f < F; ++f)
      for (int w = 0; w < W; ++w) s += rows[(((int64_t)b * F + f) * W + w) * n_acc + 3 * S];
    g.gz_sum[b] = (float)s;
  }
}

// moments[b,f,w][n_acc] -> rms, rms_field and (want_grad) gradients.  One CTA per lens.
__global__ void k_spot_finalize(const double *mom, const float *ref_y, int B, int F, int W, int S,
                                double n_rays, int want_grad, TlSpotOut out) {
  extern __shared__ double sh[];   // alpha[F], shift[F], rms[F]
  double *alpha = sh, *shift = sh + F, *rmsf = sh + 2 * F;
  const int b = blockIdx.x;
  const int n_acc = want_grad ? 6 * S + 5 : 3;
  const int m0 = want_grad ? 6 * S + 2 : 0;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    double s1 = 0.0, s2 = 0.0, n_ok = 0.0;
    for (int w = 0; w < W; ++w) {
      const double *row = mom + (((int64_t)b * F + f) * W + w) * n_acc + m0;
      s1 += row[0];
      s2 += row[1];
      n_ok += row[2];
    }
    const double y0 = (double)ref_y[b * F + f];
    // centroid over ALL rays (failed rays sit at y = 0, rtl:695-697), relative to y0
    const double mean_rel = (s1 - (n_rays - n_ok) * y0) / n_rays;
    double ss = s2 - 2.0 * mean_rel * s1 + n_ok * mean_rel * mean_rel;
    if (ss < 0.0) ss = 0.0;
    const double rms = sqrt(ss / n_rays);                            // rtl:699
    const double resid = (s1 - n_ok * mean_rel) / n_rays;            // mean of ok deviations
    rmsf[f] = rms;
    alpha[f] = 1.0 / ((double)F * n_rays * rms);
    shift[f] = mean_rel + resid;
    out.rms_field[b * F + f] = (float)rms;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int f = 0; f < F; ++f) s += rmsf[f];
    out.rms[b] = (float)(s / F);
  }
  if (!want_grad) return;
  const int per_lens = 2 * S + W * S + 1;
  for (int j = threadIdx.x; j < per_lens; j += blockDim.x) {
    double s = 0.0;
    if (j < 2 * S) {
      const int base = (j < S) ? j : 2 * S + (j - S);   // weighted slot; plain slot = +S
      for (int f = 0; f < F; ++f) {
        double a = 0.0, bsum = 0.0;
        for (int w = 0; w < W; ++w) {
          const double *row = mom + (((int64_t)b * F + f) * W + w) * n_acc;
          a += row[base];
          bsum += row[base + S];
        }
        s += alpha[f] * (a - shift[f] * bsum);
      }
      if (j < S) out.gc[b * S + j] = (float)s;
      else out.gt[b * S + (j - S)] = (float)s;
    } else if (j < 2 * S + W * S) {
      const int jj = j - 2 * S, w = jj / S, k = jj % S;
      for (int f = 0; f < F; ++f) {
        const double *row = mom + (((int64_t)b * F + f) * W + w) * n_acc;
        s += alpha[f] * (row[4 * S + k] - shift[f] * row[5 * S + k]);
      }
      out.gmu[((int64_t)b * W + w) * S + k] = (float)s;
    } else {
      for (int f = 0; f < F; ++f) {
        double a = 0.0, bsum = 0.0;
        for (int w = 0; w < W; ++w) {
          const double *row = mom + (((int64_t)b * F + f) * W + w) * n_acc;
          a += row[6 * S];
          bsum += row[6 * S + 1];
        }
        s += alpha[f] * (a - shift[f] * bsum);
      }
      out.gz[b] = (float)s;
    }
  }
}

// --------------------------------------------------------------------------
// compute_rms2d (rtl:678-702) on materialised y / ok, for every lens
// --------------------------------------------------------------------------
constexpr int kRmsThreads = 256;

// partial[(b*F+f)*nchunks + chunk][4] = {sum_all d, sum_ok d, sum_ok d^2, n_ok}, d = y - y[b,f,0,0]
__global__ void __launch_bounds__(kRmsThreads)
k_rms_partial(const float *y, const uint8_t *ok, int64_t per_field, int nchunks, int64_t chunk_len,
              double *partial) {
  const int chunk = blockIdx.x % nchunks;
  const int64_t bf = blockIdx.x / nchunks;
  const float *yy = y + bf * per_field;
  const uint8_t *kk = ok + bf * per_field;
  const float ref = yy[0];
  const int64_t lo = chunk * chunk_len, hi = min(per_field, lo + chunk_len);
  double s_all = 0.0, s1 = 0.0, s2 = 0.0, n = 0.0;
  for (int64_t base = lo; base < hi; base += (int64_t)kRmsThreads * 64) {
    float a = 0.f, b1 = 0.f, b2 = 0.f, c = 0.f;   // fp32 over <= 64 terms, then fp64
    const int64_t stop = min(hi, base + (int64_t)kRmsThreads * 64);
    for (int64_t i = base + threadIdx.x; i < stop; i += kRmsThreads) {
      const float d = yy[i] - ref;
      a += d;
      if (kk[i]) {
        b1 += d;
        b2 = ffma(d, d, b2);
        c += 1.f;
      }
    }
    s_all += a; s1 += b1; s2 += b2; n += c;
  }
  __shared__ double red[4][kRmsThreads / 32];
  double v[4] = {s_all, s1, s2, n};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double t = v[q];
    for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xffffffffu, t, d);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = t;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int q = 0; q < kRmsThreads / 32; ++q) t += red[threadIdx.x][q];
    partial[(int64_t)blockIdx.x * 4 + threadIdx.x] = t;
  }
}

// stats[b,f] = {mean (absolute), resid, alpha, rms}
__global__ void k_rms_finalize(const float *y, const double *partial, int B, int F, int64_t per_field,
                               int nchunks, float *rms, float *rms_field, double *stats) {
  extern __shared__ double sh[];
  const int b = blockIdx.x;
  const double n_rays = (double)per_field;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const int64_t bf = (int64_t)b * F + f;
    double s_all = 0.0, s1 = 0.0, s2 = 0.0, n_ok = 0.0;
    for (int c = 0; c < nchunks; ++c) {
      const double *row = partial + (bf * nchunks + c) * 4;
      s_all += row[0]; s1 += row[1]; s2 += row[2]; n_ok += row[3];
    }
    const double ref = (double)y[bf * per_field];
    const double mean_rel = s_all / n_rays;
    double ss = s2 - 2.0 * mean_rel * s1 + n_ok * mean_rel * mean_rel;
    if (ss < 0.0) ss = 0.0;
    const double r = sqrt(ss / n_rays);
    sh[f] = r;
    rms_field[bf] = (float)r;
    stats[bf * 4 + 0] = ref + mean_rel;
    stats[bf * 4 + 1] = (s1 - n_ok * mean_rel) / n_rays;
    stats[bf * 4 + 2] = 1.0 / ((double)F * n_rays * r);
    stats[bf * 4 + 3] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int f = 0; f < F; ++f) s += sh[f];
    rms[b] = (float)(s / F);
  }
}

__global__ void k_rms_bwd(const float *y, const uint8_t *ok, const double *stats,
                          const float *grad_rms, int F, int64_t per_field, int64_t total, float *gy) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t bf = i / per_field;
  const double *st = stats + bf * 4;
  const double dev = ok[i] ? ((double)y[i] - st[0]) : 0.0;
  gy[i] = (float)((double)grad_rms[bf / F] * st[2] * (dev - st[1]));
}

// --------------------------------------------------------------------------
// host-side planning
// --------------------------------------------------------------------------
int validate(const TlProblem *pb, int max_s) {
  if (!pb) return fail(TL_ERR_INVALID, "problem is NULL%s");
  if (pb->B < 1 || pb->F < 1 || pb->P < 1 || pb->W < 1 || pb->S < 1)
    return fail(TL_ERR_INVALID, "B, F, P, W, S must all be >= 1%s");
  if (pb->S > max_s) return fail(TL_ERR_INVALID, "too many surfaces for this entry point%s");
  if (!pb->x.ptr || !pb->y.ptr || !pb->z.ptr || !pb->cx.ptr || !pb->cy.ptr || !pb->c || !pb->t ||
      !pb->mu || !pb->live)
    return fail(TL_ERR_INVALID, "NULL input pointer%s");
  if (pb->arith != TL_ARITH_GUARDED && pb->arith != TL_ARITH_EXACT)
    return fail(TL_ERR_INVALID, "unknown arithmetic policy%s");
  if ((int64_t)pb->B * pb->F * pb->W > (1 << 24))
    return fail(TL_ERR_INVALID, "B*F*W too large%s");
  return TL_OK;
}

struct Plan {
  int nchunks = 1, chunk_len = 1, n_blocks = 1;
  size_t smem = 0;
};

size_t adj_smem_bytes(int S, int n_acc, bool with_state) {
  const size_t state = with_state ? (size_t)6 * S * kTraceThreads : 0;
  const size_t red = (size_t)(kTraceThreads / 32) * n_acc;
  return (5 * (size_t)S + (state > red ? state : red)) * sizeof(float);
}

// Chunks per (b,f,w): enough CTAs to fill the machine a few times over and few
// enough rays per thread that fp32 accumulators stay short.
Plan make_plan(int sms, int bfw, int n_pupil, int threads, int ctas_per_sm) {
  Plan pl;
  const int64_t want_blocks = (int64_t)sms * ctas_per_sm * 2;
  int64_t nchunks = (want_blocks + bfw - 1) / bfw;
  const int64_t min_chunks = ((int64_t)n_pupil + (int64_t)threads * kMaxRaysPerThread - 1) /
                             ((int64_t)threads * kMaxRaysPerThread);
  const int64_t max_chunks = ((int64_t)n_pupil + threads - 1) / threads;
  if (nchunks < min_chunks) nchunks = min_chunks;
  if (nchunks > max_chunks) nchunks = max_chunks;
  if (nchunks < 1) nchunks = 1;
  pl.nchunks = (int)nchunks;
  pl.chunk_len = (int)(((int64_t)n_pupil + nchunks - 1) / nchunks);
  pl.n_blocks = bfw * pl.nchunks;
  return pl;
}

int n_acc_of(int mode, int S) {
  return mode == MODE_BWD ? 3 * S + 1 : (mode == MODE_SPOT_GRAD ? 6 * S + 5 : 3);
}

template <int NS_MAX, int MODE>
int launch_adj(const TlProblem &pb, const AdjArgs &args, const Plan &pl, cudaStream_t stream) {
  auto kernel = k_trace_adj<NS_MAX, MODE>;
  if (pl.smem > 48 * 1024)
    TL_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)pl.smem));
  kernel<<<pl.n_blocks, kTraceThreads, pl.smem, stream>>>(pb, args);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

int dispatch_adj(int mode, const TlProblem &pb, const AdjArgs &args, const Plan &pl,
                 cudaStream_t stream) {
  const int S = pb.S;
  if (mode == MODE_SPOT_EVAL) return launch_adj<1, MODE_SPOT_EVAL>(pb, args, pl, stream);
  if (mode == MODE_SPOT_GRAD) {
    if (S <= 4) return launch_adj<4, MODE_SPOT_GRAD>(pb, args, pl, stream);
    if (S <= 8) return launch_adj<8, MODE_SPOT_GRAD>(pb, args, pl, stream);
    if (S <= 12) return launch_adj<12, MODE_SPOT_GRAD>(pb, args, pl, stream);
    return launch_adj<16, MODE_SPOT_GRAD>(pb, args, pl, stream);
  }
  if (S <= 4) return launch_adj<4, MODE_BWD>(pb, args, pl, stream);
  if (S <= 8) return launch_adj<8, MODE_BWD>(pb, args, pl, stream);
  if (S <= 12) return launch_adj<12, MODE_BWD>(pb, args, pl, stream);
  if (S <= 16) return launch_adj<16, MODE_BWD>(pb, args, pl, stream);
  return launch_adj<32, MODE_BWD>(pb, args, pl, stream);
}

int plan_adj(const TlProblem &pb, int mode, Plan &pl) {
  DeviceInfo info;
  int rc = device_info(info);
  if (rc) return rc;
  const int n_acc = n_acc_of(mode, pb.S);
  const size_t smem = adj_smem_bytes(pb.S, n_acc, mode != MODE_SPOT_EVAL);
  if (smem > 227 * 1024) return fail(TL_ERR_INVALID, "surface count needs too much shared memory%s");
  int per_sm = (int)((200 * 1024) / (smem > 1024 ? smem : 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  const int n_pupil = pb.p_end - pb.p_begin;
  pl = make_plan(info.sms, pb.B * pb.F * pb.W, n_pupil, kTraceThreads, per_sm);
  pl.smem = smem;
  return TL_OK;
}

size_t align8(size_t v) { return (v + 7) & ~(size_t)7; }

}  // namespace

// --------------------------------------------------------------------------
// C ABI
// --------------------------------------------------------------------------
extern "C" {

int tl_abi_version(void) { return TL_ABI_VERSION; }
const char *tl_last_error(void) { return g_error; }
int64_t tl_launch_count(void) { return (int64_t)g_launches.load(); }

int tl_trace_fwd(const TlProblem *pb, const TlTraceOut *out, void *stream_) {
  int rc = validate(pb, TL_MAX_SURFACES_FWD);
  if (rc) return rc;
  if (!out || !out->x || !out->y || !out->cx || !out->cy || !out->ok || !out->backward)
    return fail(TL_ERR_INVALID, "NULL output pointer%s");
  DeviceInfo info;
  rc = device_info(info);
  if (rc) return rc;
  const Plan pl = make_plan(info.sms, pb->B * pb->F * pb->W, pb->P, kFwdThreads, 4);
  const size_t smem = 5 * (size_t)pb->S * sizeof(float);
  k_trace_fwd<<<pl.n_blocks, kFwdThreads, smem, (cudaStream_t)stream_>>>(*pb, *out, pl.nchunks,
                                                                         pl.chunk_len);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

size_t tl_trace_bwd_workspace(const TlProblem *pb) {
  if (validate(pb, TL_MAX_SURFACES_BWD)) return 0;
  TlProblem full = *pb;
  full.p_begin = 0;
  full.p_end = pb->P;
  Plan pl;
  if (plan_adj(full, MODE_BWD, pl)) return 0;
  const size_t n_acc = n_acc_of(MODE_BWD, pb->S);
  return align8((size_t)pl.n_blocks * n_acc * sizeof(double)) +
         align8((size_t)pb->B * pb->F * pb->W * n_acc * sizeof(double));
}

int tl_trace_bwd(const TlProblem *pb_, const TlSeeds *seeds, const TlGrads *grads, void *workspace,
                 size_t workspace_bytes, void *stream_) {
  int rc = validate(pb_, TL_MAX_SURFACES_BWD);
  if (rc) return rc;
  if (!seeds || !grads || !grads->gc || !grads->gt || !grads->gmu || !grads->gz_sum)
    return fail(TL_ERR_INVALID, "NULL seeds/grads%s");
  TlProblem pb = *pb_;
  pb.p_begin = 0;
  pb.p_end = pb.P;
  Plan pl;
  rc = plan_adj(pb, MODE_BWD, pl);
  if (rc) return rc;
  const int n_acc = n_acc_of(MODE_BWD, pb.S);
  const size_t part_bytes = align8((size_t)pl.n_blocks * n_acc * sizeof(double));
  const int rows = pb.B * pb.F * pb.W;
  const size_t rows_bytes = align8((size_t)rows * n_acc * sizeof(double));
  if (!workspace || workspace_bytes < part_bytes + rows_bytes)
    return fail(TL_ERR_WORKSPACE, "workspace too small for tl_trace_bwd%s");
  cudaStream_t stream = (cudaStream_t)stream_;
  AdjArgs args;
  args.seeds = *seeds;
  args.grads = *grads;
  args.partial = (double *)workspace;
  args.ref_y = nullptr;
  args.nchunks = pl.nchunks;
  args.chunk_len = pl.chunk_len;
  args.n_acc = n_acc;
  rc = dispatch_adj(MODE_BWD, pb, args, pl, stream);
  if (rc) return rc;
  double *rowbuf = (double *)((char *)workspace + part_bytes);
  const int64_t n = (int64_t)rows * n_acc;
  k_reduce_chunks<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(args.partial, rowbuf, rows,
                                                                   pl.nchunks, n_acc);
  g_launches++;
  const int outs = pb.B * (2 * pb.S + pb.W * pb.S + 1);
  k_bwd_finalize<<<(outs + 127) / 128, 128, 0, stream>>>(rowbuf, *grads, pb.B, pb.F, pb.W, pb.S);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

size_t tl_rms_workspace(int32_t B, int32_t F, int32_t P, int32_t W) {
  if (B < 1 || F < 1 || P < 1 || W < 1) return 0;
  const int64_t per_field = (int64_t)P * W;
  int64_t nchunks = (per_field + (int64_t)kRmsThreads * 64 - 1) / ((int64_t)kRmsThreads * 64);
  if (nchunks < 1) nchunks = 1;
  if (nchunks > 4096) nchunks = 4096;
  return align8((size_t)B * F * nchunks * 4 * sizeof(double));
}

int tl_rms_fwd(const float *y, const uint8_t *ok, int32_t B, int32_t F, int32_t P, int32_t W,
               float *rms, float *rms_field, double *stats, void *workspace,
               size_t workspace_bytes, void *stream_) {
  if (!y || !ok || !rms || !rms_field || !stats || B < 1 || F < 1 || P < 1 || W < 1)
    return fail(TL_ERR_INVALID, "bad argument to tl_rms_fwd%s");
  const size_t need = tl_rms_workspace(B, F, P, W);
  if (!workspace || workspace_bytes < need)
    return fail(TL_ERR_WORKSPACE, "workspace too small for tl_rms_fwd%s");
  const int64_t per_field = (int64_t)P * W;
  const int nchunks = (int)(need / ((size_t)B * F * 4 * sizeof(double)));
  const int64_t chunk_len = (per_field + nchunks - 1) / nchunks;
  cudaStream_t stream = (cudaStream_t)stream_;
  k_rms_partial<<<B * F * nchunks, kRmsThreads, 0, stream>>>(y, ok, per_field, nchunks, chunk_len,
                                                           (double *)workspace);
  g_launches++;
  k_rms_finalize<<<B, 128, F * sizeof(double), stream>>>(y, (const double *)workspace, B, F,
                                                         per_field, nchunks, rms, rms_field, stats);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

int tl_rms_bwd(const float *y, const uint8_t *ok, const double *stats, const float *grad_rms,
               int32_t B, int32_t F, int32_t P, int32_t W, float *gy, void *stream_) {
  if (!y || !ok || !stats || !grad_rms || !gy || B < 1 || F < 1 || P < 1 || W < 1)
    return fail(TL_ERR_INVALID, "bad argument to tl_rms_bwd%s");
  const int64_t per_field = (int64_t)P * W;
  const int64_t total = per_field * B * F;
  k_rms_bwd<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      y, ok, stats, grad_rms, F, per_field, total, gy);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

int32_t tl_spot_moment_count(int32_t S, int32_t want_grad) {
  return n_acc_of(want_grad ? MODE_SPOT_GRAD : MODE_SPOT_EVAL, S);
}

size_t tl_spot_workspace(const TlProblem *pb, int32_t want_grad) {
  if (validate(pb, want_grad ? TL_MAX_SURFACES_SPOT : TL_MAX_SURFACES_FWD)) return 0;
  if (pb->p_begin < 0 || pb->p_end > pb->P || pb->p_end <= pb->p_begin) return 0;
  const int mode = want_grad ? MODE_SPOT_GRAD : MODE_SPOT_EVAL;
  Plan pl;
  if (plan_adj(*pb, mode, pl)) return 0;
  return align8((size_t)pl.n_blocks * n_acc_of(mode, pb->S) * sizeof(double));
}

int tl_spot_accumulate(const TlProblem *pb, int32_t want_grad, double *moments, float *ref_y,
                       void *workspace, size_t workspace_bytes, void *stream_) {
  int rc = validate(pb, want_grad ? TL_MAX_SURFACES_SPOT : TL_MAX_SURFACES_FWD);
  if (rc) return rc;
  if (pb->p_begin < 0 || pb->p_end > pb->P || pb->p_end <= pb->p_begin)
    return fail(TL_ERR_INVALID, "empty or out-of-range pupil slice%s");
  if (!moments || !ref_y) return fail(TL_ERR_INVALID, "NULL moments/ref_y%s");
  const int mode = want_grad ? MODE_SPOT_GRAD : MODE_SPOT_EVAL;
  Plan pl;
  rc = plan_adj(*pb, mode, pl);
  if (rc) return rc;
  const int n_acc = n_acc_of(mode, pb->S);
  if (!workspace || workspace_bytes < align8((size_t)pl.n_blocks * n_acc * sizeof(double)))
    return fail(TL_ERR_WORKSPACE, "workspace too small for tl_spot_accumulate%s");
  cudaStream_t stream = (cudaStream_t)stream_;
  AdjArgs args;
  memset(&args, 0, sizeof(args));
  args.partial = (double *)workspace;
  args.ref_y = ref_y;
  args.nchunks = pl.nchunks;
  args.chunk_len = pl.chunk_len;
  args.n_acc = n_acc;
  rc = dispatch_adj(mode, *pb, args, pl, stream);
  if (rc) return rc;
  const int rows = pb->B * pb->F * pb->W;
  const int64_t n = (int64_t)rows * n_acc;
  k_reduce_chunks<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(args.partial, moments, rows,
                                                                   pl.nchunks, n_acc);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

int tl_spot_finalize(const double *moments, const float *ref_y, int32_t B, int32_t F, int32_t W,
                     int32_t S, int64_t P_total, int32_t want_grad, const TlSpotOut *out,
                     void *stream_) {
  if (!moments || !ref_y || !out || !out->rms || !out->rms_field || B < 1 || F < 1 || W < 1 ||
      S < 1 || P_total < 1)
    return fail(TL_ERR_INVALID, "bad argument to tl_spot_finalize%s");
  if (want_grad && (!out->gc || !out->gt || !out->gmu || !out->gz))
    return fail(TL_ERR_INVALID, "NULL gradient output%s");
  if ((size_t)F * 3 * sizeof(double) > 48 * 1024)
    return fail(TL_ERR_INVALID, "too many fields%s");
  k_spot_finalize<<<B, 128, (size_t)F * 3 * sizeof(double), (cudaStream_t)stream_>>>(
      moments, ref_y, B, F, W, S, (double)P_total * (double)W, want_grad, *out);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

}  // extern "C"
