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
#include <stdlib.h>
#include <stddef.h>
#include <string.h>
#include <atomic>

#include "../../include/torchoptics_b200.h"
#include "trace_core.cuh"
#include "trace_core_asph.cuh"

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
#ifndef TL_DEFAULT_LANES_SPOT
#define TL_DEFAULT_LANES_SPOT 4
#endif
#ifndef TL_DEFAULT_LANES_BWD
#define TL_DEFAULT_LANES_BWD 4
#endif

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

// Entrance-pupil point of ray (b, f, q, w): the caller's x, y, through the ray-aiming map of
// (lens, field, wavelength) when there is one -- x_rel * x_gain, y_rel * y_gain + y_shift, clamped
// to [-2, 2] (rtl:109-112, :196-206) -- and scaled to the entrance pupil (scale_to_epd rtl:497-507).
// Every step is one individually rounded operation, like the reference's eager ops.
// (AIM = false: for the one kernel that never sees a map -- there the branch costs 27 registers)
template <bool AIM = true>
__device__ __forceinline__ void load_pupil_point(const TlProblem &pb, int b, int f, int q, int w,
                                                 float xy_scale, float &x, float &y) {
  x = pb.x.ptr[offset_of(pb.x, b, f, q, w)];
  y = pb.y.ptr[offset_of(pb.y, b, f, q, w)];
  if (AIM && pb.vig) {      // pupil vignetting first (rtl:98-104), each step individually rounded
    const float *v = pb.vig + ((int64_t)b * pb.F + f) * 3;
    x = __fmul_rn(x, v[0]);
    y = __fadd_rn(__fmul_rn(y, v[1]), v[2]);
  }
  if (AIM && pb.aim) {
    const float *a = pb.aim + (((int64_t)b * pb.F + f) * pb.W + w) * 3;
    x = fminf(fmaxf(__fmul_rn(x, a[0]), -2.0f), 2.0f);
    y = fminf(fmaxf(__fadd_rn(__fmul_rn(y, a[1]), a[2]), -2.0f), 2.0f);
  }
  if (pb.xy_scale) {
    x = __fmul_rn(x, xy_scale);
    y = __fmul_rn(y, xy_scale);
  }
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

// Parked per surface for the adjoint sweep: the hit point (x, y) and the incoming
// direction (cx, cy) (h_z and d_z are rebuilt, see trace_core.cuh).  One V-wide slot per (surface,
// component, thread): slot (k, j) of thread t is state[(k * 4 + j) * stride + t], so a
// warp's access is one contiguous, conflict-free 64/128-bit transaction per lane group.
template <bool SAVE, class V>
__device__ __forceinline__ void park(V *state, int stride, int k, V hit_x, V hit_y, V dir_x, V dir_y) {
  if (SAVE) {
    V *s = state + (size_t)k * 4 * stride;
    s[0] = hit_x;
    s[stride] = hit_y;
    s[2 * stride] = dir_x;
    s[3 * stride] = dir_y;
  }
}

// Exact-policy trace of one ray (rtl:594-675 statement by statement).  When SAVE,
// parks this ray's lane of the V-wide slots (scalar view: `stride` in floats).
// BITS: also report, as bit k of ok_bits[0], whether the ray was ok behind surface k, and as bit k
// of ok_bits[1] whether it hit surface k beyond the equator (S <= 32).
template <bool SAVE, bool BITS = false>
__device__ __noinline__ Traced trace_exact(float x, float y, float z, float cx, float cy,
                                           const Table &tab, int S, bool allow_backward,
                                           float *state, int stride, unsigned *ok_bits = nullptr) {
  Ray<float> r{x, y, z, cx, cy, exact_cz0(cx, cy)};
  bool ok = true, backward = false;
  unsigned bits = 0, flips = 0;
  for (int k = 0; k < S; ++k) {
    const float in_cx = r.cx, in_cy = r.cy;
    const Surface s{tab.c[k], tab.t[k], tab.mu[k]};
    exact_surface(r, s, k > 0 && tab.live[k - 1], allow_backward, ok, backward);
    park<SAVE, float>(state, stride, k, r.x, r.y, in_cx, in_cy);
    if (BITS && ok) bits |= 1u << (k & 31);
    if (BITS && beyond_equator(s.c, r.z, s.t)) flips |= 1u << (k & 31);
  }
  if (BITS) {
    ok_bits[0] = bits;
    ok_bits[1] = flips;      // surfaces hit beyond the equator (branch of sweep_sphere_pen)
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

// N rays per thread (V = float, f2 or f4).
template <class V>
struct TracedN {
  Ray<V> pre;
  V x, y;
  bool ok[LaneCount<V>::value], backward[LaneCount<V>::value];
  unsigned ok_bits[LaneCount<V>::value][2];   // trace_guarded<.., BITS = true> only: ok / equator bits
};

// Guarded policy: contracted (and, for f2/f4, packed) fast path for all lanes; each
// lane that was not clearly good everywhere is then re-traced alone with the exact
// policy, overwriting its lane of the parked states.
template <bool SAVE, class V, bool BITS = false>
__device__ __forceinline__ TracedN<V> trace_guarded(V x, V y, V z, V cx, V cy, const Table &tab,
                                                    int S, bool allow_backward, int arith,
                                                    V *state, int stride) {
  constexpr int N = LaneCount<V>::value;
  TracedN<V> out;
  bool clear[N];
#pragma unroll
  for (int l = 0; l < N; ++l) clear[l] = false;
  if (arith == TL_ARITH_GUARDED) {
    Ray<V> r{x, y, z, cx, cy, fast_cz0(cx, cy)};
    V min_cos2(1.0f), min_travel(3.0e38f);
    for (int k = 0; k < S; ++k) {
      const V in_cx = r.cx, in_cy = r.cy;
      V travel;
      fast_surface(r, V(tab.c[k]), V(tab.mu[k]), V(tab.mu2[k]), V(tab.t[k]), min_cos2, travel);
      park<SAVE, V>(state, stride, k, r.x, r.y, in_cx, in_cy);
      if (k > 0 && tab.live[k - 1]) min_travel = fmin2(min_travel, travel);
      // BITS: a fast-path ray carries no branch bits, so it must be clear of the equator too
      if (BITS) min_cos2 = fmin2(min_cos2, ffma(-V(tab.c[k]), r.z + V(tab.t[k]), V(1.0f)));
    }
    out.pre = r;
    const V travel = fast_image(r);
    if (tab.live[S - 1]) min_travel = fmin2(min_travel, travel);
    out.x = r.x;
    out.y = r.y;
    const V probe = (r.x + r.y) + (r.cx + r.cy);
#pragma unroll
    for (int l = 0; l < N; ++l) {
      const float band = kBandTravelRel * fmaxf(1.0f, tab.length + fabsf(lane_get(z, l)));
      clear[l] = (lane_get(min_cos2, l) > kGuard + kBandCos2) && (lane_get(min_travel, l) > band) &&
                 (fabsf(lane_get(probe, l)) < 3.0e38f);
    }
  }
#pragma unroll
  for (int l = 0; l < N; ++l) {
    out.ok[l] = true;
    out.backward[l] = false;
    if (BITS) {
      out.ok_bits[l][0] = 0xffffffffu;
      out.ok_bits[l][1] = 0u;
    }
    if (!clear[l]) {
      const Traced one = trace_exact<SAVE, BITS>(lane_get(x, l), lane_get(y, l), lane_get(z, l),
                                                 lane_get(cx, l), lane_get(cy, l), tab, S, allow_backward,
                                                 reinterpret_cast<float *>(state) + l, N * stride,
                                                 BITS ? out.ok_bits[l] : nullptr);
      lane_set(out.pre.x, l, one.pre.x);
      lane_set(out.pre.y, l, one.pre.y);
      lane_set(out.pre.z, l, one.pre.z);
      lane_set(out.pre.cx, l, one.pre.cx);
      lane_set(out.pre.cy, l, one.pre.cy);
      lane_set(out.pre.cz, l, one.pre.cz);
      lane_set(out.x, l, one.x);
      lane_set(out.y, l, one.y);
      out.ok[l] = one.ok;
      out.backward[l] = one.backward;
    }
  }
  return out;
}

// A failed ray's parked states (and `pre`) may hold anything.  Before the packed
// adjoint runs over a lane group with some dead lanes, every dead lane is given a
// live lane's states: its adjoint then stays finite and, seeded with 0, contributes
// exact zeros.
template <class V>
__device__ __noinline__ void mirror_live_lane(V *state, int stride, int S, const bool *ok,
                                              Ray<V> &pre, V &z_in, V &x_img, V &y_img) {
  constexpr int N = LaneCount<V>::value;
  int src = 0;
  for (int l = 0; l < N; ++l)
    if (ok[l]) src = l;
  float *lanes = reinterpret_cast<float *>(state);
  for (int i = 0; i < S * 4; ++i) {
    float *slot = lanes + (size_t)i * stride * N;
    const float v = slot[src];
    for (int l = 0; l < N; ++l)
      if (!ok[l]) slot[l] = v;
  }
  V *comp[9] = {&pre.x, &pre.y, &pre.z, &pre.cx, &pre.cy, &pre.cz, &z_in, &x_img, &y_img};
  for (int j = 0; j < 9; ++j) {
    const float v = lane_get(*comp[j], src);
    for (int l = 0; l < N; ++l)
      if (!ok[l]) lane_set(*comp[j], l, v);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  return v;
}

// --------------------------------------------------------------------------
// K1: forward trace (trace_skew, rtl:594-675), two rays per thread.
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
  for (int p0 = p_lo + threadIdx.x; p0 < p_hi; p0 += 2 * kFwdThreads) {
    const int p1 = p0 + kFwdThreads;
    const bool has1 = p1 < p_hi;
    const int q1 = has1 ? p1 : p0;
    const float xy_scale = pb.xy_scale ? pb.xy_scale[b] : 1.0f;
    float x0, y0, x1, y1;
    load_pupil_point<false>(pb, b, f, p0, w, xy_scale, x0, y0);     // (an aimed problem takes k_trace_fwd_pw)
    load_pupil_point<false>(pb, b, f, q1, w, xy_scale, x1, y1);
    const f2 x(x0, x1), y(y0, y1);
    const f2 z(pb.z.ptr[offset_of(pb.z, b, f, p0, w)], pb.z.ptr[offset_of(pb.z, b, f, q1, w)]);
    const f2 cx(pb.cx.ptr[offset_of(pb.cx, b, f, p0, w)], pb.cx.ptr[offset_of(pb.cx, b, f, q1, w)]);
    const f2 cy(pb.cy.ptr[offset_of(pb.cy, b, f, p0, w)], pb.cy.ptr[offset_of(pb.cy, b, f, q1, w)]);
    const TracedN<f2> tr = trace_guarded<false, f2>(x, y, z, cx, cy, tab, pb.S,
                                                    pb.allow_backward_rays != 0, pb.arith, nullptr, 0);
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      if (l == 1 && !has1) break;
      const int64_t o = (((int64_t)b * pb.F + f) * pb.P + (l ? p1 : p0)) * pb.W + w;
      out.x[o] = lane_get(tr.x, l);
      out.y[o] = lane_get(tr.y, l);
      out.cx[o] = lane_get(tr.pre.cx, l);
      out.cy[o] = lane_get(tr.pre.cy, l);
      out.ok[o] = tr.ok[l];
      out.backward[o] = tr.backward[l];
    }
  }
}

// K1b: trace_skew(aggregate=True): the forward trace that also writes the three penalty stacks
// z_RELU, theta_norm, theta_prime_norm of every surface (rtl:641-657) as [S,B,F,P,W] arrays.
// The kernel is bound by its 12 S + 18 bytes of stores per ray, not by math, so the thread map
// follows the OUTPUT layout: a CTA works on one (lens, field) and its threads run over the
// flattened (pupil, wavelength) index -- W is the innermost axis of every output, so a warp's
// stores are 128 contiguous bytes -- with the W surface tables of the lens side by side in shared
// memory.  Guarded policy: the fast path writes its stacks as it goes; a ray that was not clearly
// good is re-traced with the exact policy, which overwrites them.
template <bool STACKS>
__global__ void __launch_bounds__(kFwdThreads)
k_trace_fwd_pw(TlProblem pb, TlTraceOut out, int nchunks, int chunk_len) {
  extern __shared__ float smem[];
  int blk = blockIdx.x;
  const int chunk = blk % nchunks; blk /= nchunks;
  const int f = blk % pb.F;
  const int b = blk / pb.F;
  const int S = pb.S, W = pb.W;
  const size_t table_stride = (table_floats(S) + 3) & ~(size_t)3;
  float length = 0.f;                                         // sum |t|: the same for every wavelength
  for (int w = 0; w < W; ++w) length = load_table(smem + w * table_stride, pb, b, w).length;   // (syncs inside)
  const int64_t row_len = (int64_t)pb.P * W;                  // (p, w) flattened
  const int64_t i_lo = (int64_t)chunk * chunk_len;
  const int64_t i_hi = min(row_len, i_lo + (int64_t)chunk_len);
  const bool allow_backward = pb.allow_backward_rays != 0;
  const int64_t plane = (int64_t)pb.B * pb.F * row_len;
  const int64_t row0 = ((int64_t)b * pb.F + f) * row_len;
  const float xy_scale = pb.xy_scale ? pb.xy_scale[b] : 1.0f;
  for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += kFwdThreads) {
    const int p = (int)(i / W), w = (int)(i % W);
    Table tab;
    {
      float *base = smem + w * table_stride;
      tab.c = base;
      tab.t = base + S;
      tab.mu = base + 2 * S;
      tab.mu2 = base + 3 * S;
      tab.live = reinterpret_cast<int *>(base + 4 * S);
      tab.length = length;
    }
    float x, y;
    load_pupil_point(pb, b, f, p, w, xy_scale, x, y);
    const float z = pb.z.ptr[offset_of(pb.z, b, f, p, w)];
    const float cx = pb.cx.ptr[offset_of(pb.cx, b, f, p, w)];
    const float cy = pb.cy.ptr[offset_of(pb.cy, b, f, p, w)];
    const int64_t o = row0 + i;
    bool clear = false;
    if (pb.arith == TL_ARITH_GUARDED) {
      Ray<float> r{x, y, z, cx, cy, fast_cz0(cx, cy)};
      float min_cos2 = 1.0f, min_travel = 3.0e38f;
      for (int k = 0; k < S; ++k) {
        float travel, ci, co;
        fast_surface(r, tab.c[k], tab.mu[k], tab.mu2[k], tab.t[k], min_cos2, travel, ci, co);
        if (k > 0 && tab.live[k - 1]) min_travel = fminf(min_travel, travel);
        if (STACKS) {
          out.z_relu[k * plane + o] = r.z <= 0.0f ? 0.0f : r.z;
          out.theta[k * plane + o] = fast_angle_norm(ci);
          out.theta_prime[k * plane + o] = fast_angle_norm(co);
        }
      }
      const float pre_cx = r.cx, pre_cy = r.cy;
      const float travel = fast_image(r);
      if (tab.live[S - 1]) min_travel = fminf(min_travel, travel);
      const float band = kBandTravelRel * fmaxf(1.0f, tab.length + fabsf(z));
      clear = (min_cos2 > kGuard + kBandCos2) && (min_travel > band) &&
              (fabsf((r.x + r.y) + (pre_cx + pre_cy)) < 3.0e38f);
      if (clear) {
        out.x[o] = r.x;
        out.y[o] = r.y;
        out.cx[o] = pre_cx;
        out.cy[o] = pre_cy;
        out.ok[o] = 1;
        out.backward[o] = 0;
      }
    }
    if (!clear) {
      Ray<float> r{x, y, z, cx, cy, exact_cz0(cx, cy)};
      bool ok = true, backward = false;
      for (int k = 0; k < S; ++k) {
        const Surface s{tab.c[k], tab.t[k], tab.mu[k]};
        Penalty pen;
        exact_surface_t<STACKS>(r, s, k > 0 && tab.live[k - 1], allow_backward, ok, backward, &pen);
        if (STACKS) {
          out.z_relu[k * plane + o] = pen.z_relu;
          out.theta[k * plane + o] = pen.theta;
          out.theta_prime[k * plane + o] = pen.theta_prime;
        }
      }
      const float pre_cx = r.cx, pre_cy = r.cy;
      exact_image(r, tab.live[S - 1] != 0, allow_backward, ok, backward);
      out.x[o] = r.x;
      out.y[o] = r.y;
      out.cx[o] = pre_cx;
      out.cy[o] = pre_cy;
      out.ok[o] = ok;
      out.backward[o] = backward;
    }
  }
}

// Reference height of every (lens, field): the image height of the FIRST ray of the bundle (pupil
// point 0 -- the pupil centre, i.e. the chief ray, for the polar grids of `circle` -- wavelength 0),
// traced with the fast policy.  The spot sums are centred on it; any value near the centroid does
// (it cancels exactly in k_spot_finalize), it only has to be the same number on every CTA and every
// rank -- it depends on nothing but the prescription and the (whole, unsharded) pupil array and is
// computed by the same instruction sequence everywhere.  (A member of the bundle rather than the
// point (0, 0): a bundle that does not surround the pupil centre stays well centred too.)
// (c_row / t_row / mu_row: the lens' surface rows and its index ratios at wavelength 0, wherever the
// caller has them closest -- global memory by default)
__device__ __forceinline__ void chief_ray_one(const TlProblem &pb, int i, float *ref_y, const float *c_row = nullptr,
                                              const float *t_row = nullptr, const float *mu_row = nullptr) {
  const int b = i / pb.F, f = i % pb.F, S = pb.S;
  if (!c_row) c_row = pb.c + (int64_t)b * S;
  if (!t_row) t_row = pb.t + (int64_t)b * S;
  if (!mu_row) mu_row = pb.mu + ((int64_t)b * pb.W) * S;
  const float cx = pb.cx.ptr[offset_of(pb.cx, b, f, 0, 0)], cy = pb.cy.ptr[offset_of(pb.cy, b, f, 0, 0)];
  float x0, y0;
  load_pupil_point(pb, b, f, 0, 0, pb.xy_scale ? pb.xy_scale[b] : 1.0f, x0, y0);
  Ray<float> r{x0, y0, pb.z.ptr[offset_of(pb.z, b, f, 0, 0)], cx, cy, fast_cz0(cx, cy)};
  float min_cos2 = 1.0f, travel;
  for (int k = 0; k < S; ++k) {
    const float mu = mu_row[k];
    fast_surface(r, c_row[k], mu, mu * mu, t_row[k], min_cos2, travel);
  }
  fast_image(r);
  ref_y[i] = (min_cos2 > kGuard && fabsf(r.y) < 3.0e38f) ? r.y : 0.f;
}

__global__ void k_chief_rays(TlProblem pb, float *ref_y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < pb.B * pb.F) chief_ray_one(pb, i, ref_y);
}

// --------------------------------------------------------------------------
// K2/K3: forward + adjoint in one pass, N rays per thread (packed f32x2 math).
//   MODE_BWD        seeds come from the caller (autograd of trace_skew)
//   MODE_SPOT_GRAD  unit seed on y; accumulates both sum(J) and sum((y-y0) J) so
//                   that the RMS gradient is assembled after the reduction
//   MODE_SPOT_EVAL  forward moments only (no state, no adjoint)
//
// Persistent CTAs.  The work is the list of (row, group) items, row = (b, f, w),
// group = kTraceThreads * N consecutive pupil points; CTA i owns the contiguous
// slice [i G / n, (i + 1) G / n) of that list, so every CTA does the same amount of
// work (no tail wave) and crosses a row boundary -- where it reduces and writes its
// accumulators as one fp64 partial row ("segment") -- at most a couple of times.
// --------------------------------------------------------------------------
enum { MODE_BWD = 0, MODE_SPOT_GRAD = 1, MODE_SPOT_EVAL = 2 };

struct AdjArgs {
  TlSeeds seeds;
  TlGrads grads;
  double *partial;      // [n_blocks, max_seg, n_acc]
  const float *ref_y;   // [B,F] (spot modes)
  int groups_per_row, max_seg, n_acc;
};

#ifndef TL_ADJ_MIN_BLOCKS
#define TL_ADJ_MIN_BLOCKS 1
#endif

// PEN (MODE_BWD only): the caller also seeds the aggregate=True stacks (z_RELU, theta_norm,
// theta_prime_norm of every surface, rtl:641-657).  A ray that fails at surface j still feeds the
// stacks of the surfaces in front of j, so the sweep runs per-lane on "ok behind surface k" bits
// instead of the final ok flag, and lanes that are not ok are forced to exact zeros (no mirroring).
// PEN = PEN_SUM: no seed arrays -- every term of every ray is seeded with 1 and the kernel also sums
// the terms themselves (slot 3S+1 of the row): value and gradient of the penalty sum(Q) of
// compute_loss_out (optics_simulator_lite.py:430-450) in one pass that materialises nothing.
enum { PEN_NONE = 0, PEN_SEEDED = 1, PEN_SUM = 2 };
template <int NS_MAX, int MODE, class V, int PEN = PEN_NONE>
__global__ void __launch_bounds__(kTraceThreads, TL_ADJ_MIN_BLOCKS)
k_trace_adj(TlProblem pb, AdjArgs args) {
  extern __shared__ float smem[];
  constexpr int N = LaneCount<V>::value;
  constexpr bool kSpot = MODE != MODE_BWD;
  constexpr bool kAdjoint = MODE != MODE_SPOT_EVAL;
  constexpr int NA = kAdjoint ? NS_MAX : 1;
  constexpr int kWarps = kTraceThreads / 32;
  const int S = pb.S;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int stride = kTraceThreads;
  const bool allow_backward = pb.allow_backward_rays != 0;
  const int n_acc = args.n_acc;
  // parked states start on a 16-byte boundary behind the table
  V *state = reinterpret_cast<V *>(smem + ((table_floats(S) + 3) & ~(size_t)3)) + tid;
  float *red = smem + table_floats(S);

  const int64_t total = (int64_t)pb.B * pb.F * pb.W * args.groups_per_row;
  const int64_t g_begin = total * blockIdx.x / gridDim.x;
  const int64_t g_end = total * (blockIdx.x + 1) / gridDim.x;

  // accumulators (registers: every index below is a compile-time constant)
  float acc_c[NA], acc_t[NA], acc_mu[NA];      // sum J        (BWD: sum of gradients)
  float wac_c[NA], wac_t[NA], wac_mu[NA];      // sum (y-y0) J (SPOT_GRAD only)
  float acc_z = 0.f, wac_z = 0.f;
  float m_s1 = 0.f, m_s2 = 0.f, m_n = 0.f, m_pen = 0.f;
  Table tab;
  float y0 = 0.f, xy_scale = 1.0f;
  int row = -1, seg = 0, b = 0, f = 0, w = 0;

  auto flush = [&]() {
    // CTA reduction: warp shuffles -> smem [warp][slot] -> one fp64 partial row
    __syncthreads();
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
      if (PEN == PEN_SUM) put(3 * S + 1, m_pen);
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
    double *dst = args.partial + ((int64_t)blockIdx.x * args.max_seg + seg) * n_acc;
    for (int i = tid; i < n_acc; i += kTraceThreads) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < kWarps; ++q) s += (double)red[q * n_acc + i];
      dst[i] = s;
    }
    ++seg;
    __syncthreads();
  };

  for (int64_t g = g_begin; g < g_end; ++g) {
    const int r = (int)(g / args.groups_per_row);
    const int j = (int)(g % args.groups_per_row);
    if (r != row) {
      if (row >= 0) flush();
      row = r;
      w = r % pb.W;
      f = (r / pb.W) % pb.F;
      b = r / (pb.W * pb.F);
      tab = load_table(smem, pb, b, w);
      xy_scale = pb.xy_scale ? pb.xy_scale[b] : 1.0f;
      if (kSpot) y0 = args.ref_y[b * pb.F + f];
#pragma unroll
      for (int k = 0; k < NA; ++k) {
        acc_c[k] = acc_t[k] = acc_mu[k] = 0.f;
        wac_c[k] = wac_t[k] = wac_mu[k] = 0.f;
      }
      acc_z = wac_z = 0.f;
      m_s1 = m_s2 = m_n = m_pen = 0.f;
    }
    // lane l of this thread = pupil point p_base + l * threads + tid
    const int p_base = pb.p_begin + j * (kTraceThreads * N) + tid;
    bool has[N];
    int64_t o[N];
    V x, y, z, cx, cy;
#pragma unroll
    for (int l = 0; l < N; ++l) {
      const int p = p_base + l * kTraceThreads;
      has[l] = p < pb.p_end;
      const int q = has[l] ? p : p_base;          // past the end: a copy of lane 0 (p_base is valid)
      o[l] = (((int64_t)b * pb.F + f) * pb.P + q) * pb.W + w;
      float px, py;
      load_pupil_point(pb, b, f, q, w, xy_scale, px, py);
      lane_set(x, l, px);
      lane_set(y, l, py);
      lane_set(z, l, pb.z.ptr[offset_of(pb.z, b, f, q, w)]);
      lane_set(cx, l, pb.cx.ptr[offset_of(pb.cx, b, f, q, w)]);
      lane_set(cy, l, pb.cy.ptr[offset_of(pb.cy, b, f, q, w)]);
    }
    if (p_base >= pb.p_end) continue;   // whole thread past the end of the row
    TracedN<V> tr = trace_guarded<kAdjoint, V, PEN != PEN_NONE>(x, y, z, cx, cy, tab, S, allow_backward, pb.arith,
                                                    state, stride);
    bool live[N], any_live = false, all_ok = true;
    V alive, wgt(0.f);
#pragma unroll
    for (int l = 0; l < N; ++l) {
      live[l] = tr.ok[l] && has[l];
      any_live = any_live || live[l];
      all_ok = all_ok && tr.ok[l];
      lane_set(alive, l, live[l] ? 1.0f : 0.0f);
    }
    if (kSpot) {
      wgt = (tr.y - V(y0)) * alive;
#pragma unroll
      for (int l = 0; l < N; ++l)
        if (!live[l]) lane_set(wgt, l, 0.f);      // a dead lane's y may be anything
      m_s1 += lane_sum(wgt);
      m_s2 = lane_dot(wgt, wgt, m_s2);
      m_n += lane_sum(alive);
    }
    if (kAdjoint) {
      Ray<V> a{V(0.f), V(0.f), V(0.f), V(0.f), V(0.f), V(0.f)};
      unsigned bits[N], flips[N];
      bool any_bits = false;
      if (PEN != PEN_NONE) {
#pragma unroll
        for (int l = 0; l < N; ++l) {
          bits[l] = has[l] ? tr.ok_bits[l][0] : 0u;
          flips[l] = tr.ok_bits[l][1];
          any_bits = any_bits || bits[l] != 0u || (has[l] && (PEN == PEN_SUM || args.seeds.gz_relu != nullptr));
        }
      }
      if (any_live || (PEN != PEN_NONE && any_bits)) {
        if (PEN == PEN_NONE && !all_ok) mirror_live_lane<V>(state, stride, S, tr.ok, tr.pre, z, tr.x, tr.y);
        V sx(0.f), sy(0.f), scx(0.f), scy(0.f);
        if (MODE == MODE_SPOT_GRAD) {
          sy = alive;
        } else {
#pragma unroll
          for (int l = 0; l < N; ++l) {
            if (!live[l] || PEN == PEN_SUM) continue;
            if (args.seeds.gx) lane_set(sx, l, args.seeds.gx[o[l]]);
            if (args.seeds.gy) lane_set(sy, l, args.seeds.gy[o[l]]);
            if (args.seeds.gcx) lane_set(scx, l, args.seeds.gcx[o[l]]);
            if (args.seeds.gcy) lane_set(scy, l, args.seeds.gcy[o[l]]);
          }
        }
        Sweep<V> sw = sweep_begin(tr.pre, tr.x, tr.y, sx, sy, scx, scy);
#pragma unroll
        for (int k = NS_MAX - 1; k >= 0; --k) {
          if (k >= S) continue;
          const V *slot = state + (size_t)k * 4 * stride;
          SurfaceGrad<V> g;
          if constexpr (PEN != PEN_NONE) {
            // seeds of this surface's stacks ([S,B,F,P,W]); only lanes ok behind surface k carry
            // angle seeds, and their z_RELU seed enters through the hit point
            const int64_t plane = (int64_t)pb.B * pb.F * pb.P * pb.W;
            V pz(0.f), pth(0.f), pthp(0.f), branch(1.0f);
            float dead_gt[N];
#pragma unroll
            for (int l = 0; l < N; ++l) {
              dead_gt[l] = 0.f;
              if ((flips[l] >> k) & 1u) lane_set(branch, l, -1.0f);
              if (!has[l]) continue;
              const int64_t at = (int64_t)k * plane + o[l];
              if (PEN == PEN_SUM) {
                if ((bits[l] >> k) & 1u) {
                  lane_set(pz, l, 1.0f);
                  lane_set(pth, l, 1.0f);
                  lane_set(pthp, l, 1.0f);
                } else {      // failed ray: theta = theta' = 1, z = 0 - t[k] (rtl:639, :653-654)
                  const float z_dead = -tab.t[k];
                  m_pen += 2.0f + fmaxf(z_dead, 0.0f);
                  if (z_dead > 0.f) dead_gt[l] = -1.0f;
                }
              } else if ((bits[l] >> k) & 1u) {
                if (args.seeds.gz_relu) lane_set(pz, l, args.seeds.gz_relu[at]);
                if (args.seeds.gtheta) lane_set(pth, l, args.seeds.gtheta[at]);
                if (args.seeds.gtheta_prime) lane_set(pthp, l, args.seeds.gtheta_prime[at]);
              } else if (args.seeds.gz_relu && -tab.t[k] > 0.f) {
                dead_gt[l] = -args.seeds.gz_relu[at];     // a failed ray sits at z = 0 - t[k] (rtl:639)
              }
            }
            V cos_in, cos_out, z_behind;
            g = sweep_sphere_pen(sw, slot[0], slot[stride], slot[2 * stride], slot[3 * stride],
                                 V(tab.c[k]), V(tab.t[k]), V(tab.mu[k]), V(tab.mu2[k]), pz, pth, pthp, branch,
                                 cos_in, cos_out, z_behind);
            if (PEN == PEN_SUM) {
#pragma unroll
              for (int l = 0; l < N; ++l)
                if (has[l] && ((bits[l] >> k) & 1u))
                  m_pen += (fast_angle_norm(lane_get(cos_in, l)) + fast_angle_norm(lane_get(cos_out, l))) +
                           fmaxf(lane_get(z_behind, l), 0.0f);
            }
#pragma unroll
            for (int l = 0; l < N; ++l) {
              if ((bits[l] >> k) & 1u) continue;
              lane_set(g.c, l, 0.f);
              lane_set(g.mu, l, 0.f);
              lane_set(g.t, l, dead_gt[l]);
              lane_set(sw.gr.x, l, 0.f); lane_set(sw.gr.y, l, 0.f); lane_set(sw.gr.z, l, 0.f);
              lane_set(sw.gd.x, l, 0.f); lane_set(sw.gd.y, l, 0.f); lane_set(sw.gd.z, l, 0.f);
            }
          } else {
            g = sweep_sphere(sw, slot[0], slot[stride], slot[2 * stride], slot[3 * stride],
                             V(tab.c[k]), V(tab.t[k]), V(tab.mu[k]), V(tab.mu2[k]));
          }
          acc_c[k] += lane_sum(g.c);
          acc_t[k] += lane_sum(g.t);
          acc_mu[k] += lane_sum(g.mu);
          if (MODE == MODE_SPOT_GRAD) {
            wac_c[k] = lane_dot(wgt, g.c, wac_c[k]);
            wac_t[k] = lane_dot(wgt, g.t, wac_t[k]);
            wac_mu[k] = lane_dot(wgt, g.mu, wac_mu[k]);
          }
        }
        sweep_end(sw, z, a.x, a.y, a.z, a.cx, a.cy);
        acc_z += lane_sum(a.z);
        if (MODE == MODE_SPOT_GRAD) wac_z = lane_dot(wgt, a.z, wac_z);
      }
      if (MODE == MODE_BWD) {
#pragma unroll
        for (int l = 0; l < N; ++l) {
          if (!has[l]) continue;
          if (args.grads.gx) args.grads.gx[o[l]] = lane_get(a.x, l) * xy_scale;
          if (args.grads.gy) args.grads.gy[o[l]] = lane_get(a.y, l) * xy_scale;
          if (args.grads.gz) args.grads.gz[o[l]] = lane_get(a.z, l);
          if (args.grads.gcx) args.grads.gcx[o[l]] = lane_get(a.cx, l);
          if (args.grads.gcy) args.grads.gcy[o[l]] = lane_get(a.cy, l);
        }
      }
    }
  }
  if (row >= 0) flush();
}

// CTA that owns work item g of `total` when they are dealt as contiguous slices
__device__ __forceinline__ int64_t owner_of(int64_t g, int64_t total, int64_t n_blocks) {
  return ((g + 1) * n_blocks - 1) / total;
}

// partial[block, segment, n_acc] -> dst[row, n_acc], fixed summation order
__global__ void k_reduce_rows(const double *partial, double *dst, int n_rows, int groups_per_row,
                              int n_blocks, int max_seg, int n_acc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_rows * n_acc) return;
  const int row = (int)(i / n_acc), slot = (int)(i % n_acc);
  const int64_t total = (int64_t)n_rows * groups_per_row;
  const int64_t first = owner_of((int64_t)row * groups_per_row, total, n_blocks);
  const int64_t last = owner_of((int64_t)(row + 1) * groups_per_row - 1, total, n_blocks);
  double s = 0.0;
  for (int64_t blk = first; blk <= last; ++blk) {
    const int64_t lo = total * blk / n_blocks, hi = total * (blk + 1) / n_blocks;
    if (hi <= lo) continue;                      // an owner with an empty slice wrote nothing
    const int64_t first_row = lo / groups_per_row;
    s += partial[((blk * max_seg) + (row - first_row)) * n_acc + slot];
  }
  dst[i] = s;
}

#include "trace_kernels_gen.cuh"

// --------------------------------------------------------------------------
// K3b: the fused spot pass for MANY SMALL rows -- the reference's real workload shape (SURVEY
// section 8f-3: hundreds of lenses, 8 fields x 3 wavelengths x 8x8 pupil = 64 rays per row).
// k_trace_adj gives a whole CTA (128 threads x N rays) to one (lens, field, wavelength) row and
// pays a CTA-wide table load, two barriers and a 71-value block reduction per row: with 64 rays
// per row most lanes idle and the per-row overhead dominates (measured 15 G events/s).
// Here a WARP owns a row: its own surface table and parked states in shared memory, no
// __syncthreads anywhere, and the row's sums leave through a halving-butterfly transpose
// (warp_transpose_sum: 31 shuffles per 32 values) straight into `moments` -- no partial rows,
// no k_reduce_rows.  Rows are dealt to the warps of a persistent grid round-robin.
// Same arithmetic (trace_guarded, sweep_sphere), same moment layout as k_trace_adj.
// --------------------------------------------------------------------------
__device__ __forceinline__ Table load_table_warp(float *base, const TlProblem &pb, int b, int w, int lane) {
  Table tab;
  const int S = pb.S;
  tab.c = base;
  tab.t = base + S;
  tab.mu = base + 2 * S;
  tab.mu2 = base + 3 * S;
  tab.live = reinterpret_cast<int *>(base + 4 * S);
  __syncwarp();                                   // the previous row's readers are done
  for (int k = lane; k < S; k += 32) {
    const float m = pb.mu[((int64_t)b * pb.W + w) * S + k];
    tab.c[k] = pb.c[(int64_t)b * S + k];
    tab.t[k] = pb.t[(int64_t)b * S + k];
    tab.mu[k] = m;
    tab.mu2[k] = m * m;
    tab.live[k] = pb.live[(int64_t)b * S + k] != 0;
  }
  __syncwarp();
  float len = 0.f;
  for (int k = 0; k < S; ++k) len += fabsf(tab.t[k]);
  tab.length = len;
  return tab;
}

template <int NS_MAX, bool GRAD, class V>
__global__ void __launch_bounds__(kTraceThreads)
k_spot_rows(TlProblem pb, const float *ref_y, double *moments, int n_acc) {
  extern __shared__ float smem[];
  constexpr int N = LaneCount<V>::value;
  constexpr int NA = GRAD ? NS_MAX : 1;
  constexpr int kWarps = kTraceThreads / 32;
  const int S = pb.S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool allow_backward = pb.allow_backward_rays != 0;
  const size_t tab_floats = (table_floats(S) + 3) & ~(size_t)3;
  const size_t state_floats = GRAD ? (size_t)4 * S * 32 * N : 0;
  float *wbase = smem + warp * (tab_floats + state_floats);
  V *state = reinterpret_cast<V *>(wbase + tab_floats) + lane;
  constexpr int stride = 32;
  const int n_rows = pb.B * pb.F * pb.W;
  const int groups = (pb.p_end - pb.p_begin + 32 * N - 1) / (32 * N);

  for (int row = blockIdx.x * kWarps + warp; row < n_rows; row += gridDim.x * kWarps) {
    const int w = row % pb.W;
    const int f = (row / pb.W) % pb.F;
    const int b = row / (pb.W * pb.F);
    const Table tab = load_table_warp(wbase, pb, b, w, lane);
    const float xy_scale = pb.xy_scale ? pb.xy_scale[b] : 1.0f;
    const float y0 = ref_y[b * pb.F + f];
    float acc_c[NA], acc_t[NA], acc_mu[NA], wac_c[NA], wac_t[NA], wac_mu[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      acc_c[k] = acc_t[k] = acc_mu[k] = 0.f;
      wac_c[k] = wac_t[k] = wac_mu[k] = 0.f;
    }
    float acc_z = 0.f, wac_z = 0.f, m_s1 = 0.f, m_s2 = 0.f, m_n = 0.f;

    for (int j = 0; j < groups; ++j) {
      const int p_base = pb.p_begin + j * (32 * N) + lane;
      if (p_base >= pb.p_end) continue;            // whole thread past the end of the row
      bool has[N];
      V x, y, z, cx, cy;
#pragma unroll
      for (int l = 0; l < N; ++l) {
        const int p = p_base + l * 32;
        has[l] = p < pb.p_end;
        const int q = has[l] ? p : p_base;
        float px, py;
        load_pupil_point(pb, b, f, q, w, xy_scale, px, py);
        lane_set(x, l, px);
        lane_set(y, l, py);
        lane_set(z, l, pb.z.ptr[offset_of(pb.z, b, f, q, w)]);
        lane_set(cx, l, pb.cx.ptr[offset_of(pb.cx, b, f, q, w)]);
        lane_set(cy, l, pb.cy.ptr[offset_of(pb.cy, b, f, q, w)]);
      }
      TracedN<V> tr = trace_guarded<GRAD, V>(x, y, z, cx, cy, tab, S, allow_backward, pb.arith, state, stride);
      bool live[N], any_live = false, all_ok = true;
      V alive;
#pragma unroll
      for (int l = 0; l < N; ++l) {
        live[l] = tr.ok[l] && has[l];
        any_live = any_live || live[l];
        all_ok = all_ok && tr.ok[l];
        lane_set(alive, l, live[l] ? 1.0f : 0.0f);
      }
      V wgt = (tr.y - V(y0)) * alive;
#pragma unroll
      for (int l = 0; l < N; ++l)
        if (!live[l]) lane_set(wgt, l, 0.f);
      m_s1 += lane_sum(wgt);
      m_s2 = lane_dot(wgt, wgt, m_s2);
      m_n += lane_sum(alive);
      if (GRAD && any_live) {
        if (!all_ok) mirror_live_lane<V>(state, stride, S, tr.ok, tr.pre, z, tr.x, tr.y);
        Sweep<V> sw = sweep_begin(tr.pre, tr.x, tr.y, V(0.f), alive, V(0.f), V(0.f));
#pragma unroll
        for (int k = NS_MAX - 1; k >= 0; --k) {
          if (k >= S) continue;
          const V *slot = state + (size_t)k * 4 * stride;
          const SurfaceGrad<V> g = sweep_sphere(sw, slot[0], slot[stride], slot[2 * stride], slot[3 * stride],
                                                V(tab.c[k]), V(tab.t[k]), V(tab.mu[k]), V(tab.mu2[k]));
          acc_c[k] += lane_sum(g.c);
          acc_t[k] += lane_sum(g.t);
          acc_mu[k] += lane_sum(g.mu);
          wac_c[k] = lane_dot(wgt, g.c, wac_c[k]);
          wac_t[k] = lane_dot(wgt, g.t, wac_t[k]);
          wac_mu[k] = lane_dot(wgt, g.mu, wac_mu[k]);
        }
        V ax, ay, az, acx, acy;
        sweep_end(sw, z, ax, ay, az, acx, acy);
        acc_z += lane_sum(az);
        wac_z = lane_dot(wgt, az, wac_z);
      }
    }

    // the row's sums: 6 values per surface go through the transpose five surfaces at a time
    // (lane 6 q + typ of batch i ends with the warp total of value typ of surface 5 i + q: slot
    // typ * S + k of the row, exactly k_trace_adj's layout), the five scalars through plain sums
    double *dst = moments + (int64_t)row * n_acc;
    if (GRAD) {
      constexpr int kPerBatch = 5;
#pragma unroll
      for (int i = 0; i < (NS_MAX + kPerBatch - 1) / kPerBatch; ++i) {
        float v[32];
#pragma unroll
        for (int q = 0; q < kPerBatch; ++q) {
          const int k = i * kPerBatch + q;
          const bool in = k < NA;
          v[6 * q + 0] = in ? wac_c[k < NA ? k : 0] : 0.f;
          v[6 * q + 1] = in ? acc_c[k < NA ? k : 0] : 0.f;
          v[6 * q + 2] = in ? wac_t[k < NA ? k : 0] : 0.f;
          v[6 * q + 3] = in ? acc_t[k < NA ? k : 0] : 0.f;
          v[6 * q + 4] = in ? wac_mu[k < NA ? k : 0] : 0.f;
          v[6 * q + 5] = in ? acc_mu[k < NA ? k : 0] : 0.f;
        }
        v[30] = v[31] = 0.f;
        const float total = warp_transpose_sum(v, lane);
        const int k = i * kPerBatch + lane / 6, typ = lane % 6;
        if (lane < 30 && k < S) dst[typ * S + k] = (double)total;
      }
      const float t0 = warp_sum(wac_z), t1 = warp_sum(acc_z), t2 = warp_sum(m_s1), t3 = warp_sum(m_s2),
                  t4 = warp_sum(m_n);
      if (lane == 0) {
        dst[6 * S] = (double)t0;
        dst[6 * S + 1] = (double)t1;
        dst[6 * S + 2] = (double)t2;
        dst[6 * S + 3] = (double)t3;
        dst[6 * S + 4] = (double)t4;
      }
    } else {
      const float t2 = warp_sum(m_s1), t3 = warp_sum(m_s2), t4 = warp_sum(m_n);
      if (lane == 0) {
        dst[0] = (double)t2;
        dst[1] = (double)t3;
        dst[2] = (double)t4;
      }
    }
  }
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

// rows[b,f,w][3S+2] of the penalty pass -> penalty[b] and its gradients, all times `scale`
// (= 1 / numSequence of compute_loss_out)
__global__ void k_penalty_finalize(const double *rows, TlPenaltyOut out, int B, int F, int W, int S,
                                   double scale) {
  const int n_acc = 3 * S + 2;
  const int per_lens = 2 * S + W * S + 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * per_lens) return;
  const int b = i / per_lens;
  int j = i % per_lens;
  double s = 0.0;
  if (j < 2 * S) {                       // c (j < S) or t
    for (int f = 0; f < F; ++f)
      for (int w = 0; w < W; ++w) s += rows[(((int64_t)b * F + f) * W + w) * n_acc + j];
    if (j < S) out.gc[b * S + j] = (float)(s * scale);
    else out.gt[b * S + (j - S)] = (float)(s * scale);
  } else if (j < 2 * S + W * S) {
    j -= 2 * S;
    const int w = j / S, k = j % S;
    for (int f = 0; f < F; ++f) s += rows[(((int64_t)b * F + f) * W + w) * n_acc + 2 * S + k];
    out.gmu[((int64_t)b * W + w) * S + k] = (float)(s * scale);
  } else {
    const int slot = (j == 2 * S + W * S) ? 3 * S : 3 * S + 1;
    for (int f = 0; f < F; ++f)
      for (int w = 0; w < W; ++w) s += rows[(((int64_t)b * F + f) * W + w) * n_acc + slot];
    if (slot == 3 * S) out.gz[b] = (float)(s * scale);
    else out.penalty[b] = (float)(s * scale);
  }
}

// moments[b,f,w][n_acc] -> rms, rms_field and (want_grad) gradients.  One CTA per lens.
// The lens' rows are first staged in shared memory with coalesced loads (STAGED), so the
// dependent fp64 sums below do not each wait for a global-memory round trip.
template <bool STAGED>
__device__ __forceinline__ void spot_finalize_lens(const double *mom_global, const float *ref_y, int B, int F, int W,
                                                   int S, double n_rays, int want_grad, const TlSpotOut &out) {
  extern __shared__ double sh[];   // alpha[F], shift[F], rms[F], then (STAGED) the rows
  double *alpha = sh, *shift = sh + F, *rmsf = sh + 2 * F;
  const int b = blockIdx.x;
  const int n_acc = want_grad ? 6 * S + 5 : 3;
  const int m0 = want_grad ? 6 * S + 2 : 0;
  const double *mom = mom_global + (int64_t)b * F * W * n_acc;    // this lens' rows
  if (STAGED) {
    double *rows = sh + 3 * F;
    const int n = F * W * n_acc, bd = blockDim.x;
    int i = threadIdx.x;
    for (; i + 7 * bd < n; i += 8 * bd) {      // eight loads in flight per thread: one round trip, not eight
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = mom[i + u * bd];
#pragma unroll
      for (int u = 0; u < 8; ++u) rows[i + u * bd] = v[u];
    }
    for (; i < n; i += bd) rows[i] = mom[i];
    __syncthreads();
    mom = rows;
  }
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    double s1 = 0.0, s2 = 0.0, n_ok = 0.0;
    for (int w = 0; w < W; ++w) {
      const double *row = mom + ((int64_t)f * W + w) * n_acc + m0;
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
          const double *row = mom + ((int64_t)f * W + w) * n_acc;
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
        const double *row = mom + ((int64_t)f * W + w) * n_acc;
        s += alpha[f] * (row[4 * S + k] - shift[f] * row[5 * S + k]);
      }
      out.gmu[((int64_t)b * W + w) * S + k] = (float)s;
    } else {
      for (int f = 0; f < F; ++f) {
        double a = 0.0, bsum = 0.0;
        for (int w = 0; w < W; ++w) {
          const double *row = mom + ((int64_t)f * W + w) * n_acc;
          a += row[6 * S];
          bsum += row[6 * S + 1];
        }
        s += alpha[f] * (a - shift[f] * bsum);
      }
      out.gz[b] = (float)s;
    }
  }
}

template <bool STAGED>
__global__ void k_spot_finalize(const double *mom_global, const float *ref_y, int B, int F, int W,
                                int S, double n_rays, int want_grad, TlSpotOut out) {
  spot_finalize_lens<STAGED>(mom_global, ref_y, B, F, W, S, n_rays, want_grad, out);
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
// Ray-set staging (RayTracer.trace_rays rtl:80-124) and its chain rule.
// One CTA per lens; everything is O(L * W) scalars.
// --------------------------------------------------------------------------
constexpr int kStageMaxSurfaces = 64;
constexpr double kLineC = 656.3, kLineD = 587.6, kLineF = 486.1;   // lens_modeling.py:362-364

// n(lambda) of slot s (1 for air / padding), lens_modeling.py:355-374
__device__ __forceinline__ float index_at(const TlLens &ln, int b, int s, float wl, float *dn_dnd,
                                          float *dn_dv) {
  const int64_t i = (int64_t)b * ln.L + s;
  if (dn_dnd) *dn_dnd = 0.f;
  if (dn_dv) *dn_dv = 0.f;
  if (!ln.mask_g[i]) return 1.0f;
  const float nd = ln.nd[i], v = ln.v[i];
  if (v == 0.f) {                       // dispersion-free glass
    if (dn_dnd) *dn_dnd = 1.f;
    return nd;
  }
  // constants evaluated in double and rounded once, like the Python scalars of the host code
  const float kk = (float)(1.0 / ((double)kLineF * kLineF) - 1.0 / ((double)kLineC * kLineC));
  const float ld2 = (float)((double)kLineD * kLineD);
  const float slope = (nd - 1.0f) / (v * kk);
  const float offset = nd - slope / ld2;
  const float span = 1.0f / (wl * wl) - 1.0f / ld2;
  if (dn_dnd) *dn_dnd = 1.0f + span / (v * kk);
  if (dn_dv) *dn_dv = -slope * span / v;
  return offset + slope / (wl * wl);
}

__device__ __forceinline__ float nd_at(const TlLens &ln, int b, int s) {
  const int64_t i = (int64_t)b * ln.L + s;
  return ln.mask_g[i] ? ln.nd[i] : 1.0f;
}

// ABCD matrix of slot s in front of the stop (identity for padding), rtl:314-327
struct Abcd {
  float a, b, c, d;
};
__device__ __forceinline__ Abcd slot_matrix(const TlLens &ln, int b, int s, float *ratio_out,
                                            float *power_out) {
  const int64_t i = (int64_t)b * ln.L + s;
  const bool on = ln.mask[i] && s < ln.stop_idx[b];
  const float n_in = (s == 0) ? 1.0f : (on ? nd_at(ln, b, s - 1) : 1.0f);
  const float n_out = on ? nd_at(ln, b, s) : 1.0f;
  const float cc = on ? ln.c[i] : 0.f, tt = on ? ln.t[i] : 0.f;
  const float ratio = n_in / n_out, power = cc * (ratio - 1.0f);
  if (ratio_out) *ratio_out = ratio;
  if (power_out) *power_out = power;
  return Abcd{1.0f + power * tt, ratio * tt, power, ratio};
}

// The small kernels around the fused pass are single-CTA chains of dependent loads: with the lens in
// global memory every step of a serial loop pays a ~0.6 us round trip (k_stage_ref 8 us, the finalize
// kernel 21 us of a 270 us step).  This copies lens b -- c, t, nd, v, the two masks, its stop index,
// field angle and pupil diameter, and the wavelength list -- into shared memory in ONE parallel round
// trip and returns a view of it: a TlLens whose pointers are biased so that the usual [b * L + s] /
// [b] / [w] indexing lands in the copy.  Ends with a barrier.
struct LensCopy {
  float f[4 * kStageMaxSurfaces + 2 + 16];      // c, t, nd, v, hfov, epd, wavelengths (first 16)
  uint8_t m[2 * kStageMaxSurfaces];
  int stop;
};

__device__ __forceinline__ TlLens lens_in_shared(const TlLens &ln, int b, LensCopy &buf) {
  const int L = ln.L;
  for (int i = threadIdx.x; i < 4 * L; i += blockDim.x) {
    const int which = i / L, s_ = i % L;
    const float *src = which == 0 ? ln.c : which == 1 ? ln.t : which == 2 ? ln.nd : ln.v;
    buf.f[i] = src[(int64_t)b * L + s_];
  }
  for (int i = threadIdx.x; i < 2 * L; i += blockDim.x)
    buf.m[i] = (i < L ? ln.mask : ln.mask_g)[(int64_t)b * L + i % L];
  if (threadIdx.x == 0) {
    buf.stop = ln.stop_idx[b];
    buf.f[4 * L] = ln.hfov[b];
    buf.f[4 * L + 1] = ln.epd[b];
  }
  for (int w = threadIdx.x; w < ln.W && w < 16; w += blockDim.x) buf.f[4 * L + 2 + w] = ln.wavelengths[w];
  __syncthreads();
  TlLens v = ln;
  const int64_t bias = (int64_t)b * L;
  v.c = buf.f - bias;
  v.t = buf.f + L - bias;
  v.nd = buf.f + 2 * L - bias;
  v.v = buf.f + 3 * L - bias;
  v.mask = buf.m - bias;
  v.mask_g = buf.m + L - bias;
  v.stop_idx = &buf.stop - b;
  v.hfov = buf.f + 4 * L - b;
  v.epd = buf.f + 4 * L + 1 - b;
  if (ln.W <= 16) v.wavelengths = buf.f + 4 * L + 2;
  return v;
}

__device__ __forceinline__ void stage_fwd_lens(const TlLens &ln, int b, float *mu, float *z, float *cy,
                                               float *half_epd, float *mu0_shared = nullptr) {
  if (threadIdx.x == 1 % blockDim.x) half_epd[b] = ln.epd[b] * 0.5f;
  for (int i = threadIdx.x; i < ln.W * ln.L; i += blockDim.x) {
    const int w = i / ln.L, s = i % ln.L;
    const float wl = ln.wavelengths[w];
    const float n_in = (s == 0) ? 1.0f : index_at(ln, b, s - 1, wl, nullptr, nullptr);
    const float ratio = n_in / index_at(ln, b, s, wl, nullptr, nullptr);
    mu[((int64_t)b * ln.W + w) * ln.L + s] = ratio;
    if (mu0_shared && w == 0) mu0_shared[s] = ratio;      // (the reference-height rays run at wavelength 0)
  }
  for (int f = threadIdx.x; f < ln.F; f += blockDim.x)
    cy[(int64_t)b * ln.F + f] = sinf(ln.hfov[b] * ln.rel_fields[f]);
  if (threadIdx.x == 0) {
    Abcd m{1.f, 0.f, 0.f, 1.f};
    const int n_front = min(ln.stop_idx[b], ln.L);
    for (int s = 0; s < n_front; ++s) {                // M <- M_s M
      const Abcd q = slot_matrix(ln, b, s, nullptr, nullptr);
      m = Abcd{q.a * m.a + q.b * m.c, q.a * m.b + q.b * m.d, q.c * m.a + q.d * m.c,
               q.c * m.b + q.d * m.d};
    }
    z[b] = n_front > 0 ? m.b / m.a : 0.f;
  }
}

__global__ void k_stage_fwd(TlLens ln, float *mu, float *z, float *cy, float *half_epd) {
  stage_fwd_lens(ln, blockIdx.x, mu, z, cy, half_epd);
}

// Ray aiming on the device (RayTracer.ray_aiming, rtl:129-208, one iteration, 'real' stop radius).
// One thread per (lens, field, wavelength): trace the on-axis marginal ray at the d line to the stop
// (the stop radius rs, compute_pupil_radius rtl:834-844), then the three 'tee' rays (lower / upper
// meridional, sagittal; rtl:353-360) through the surfaces in front of the stop in forward mode (D2:
// value + derivatives along the pupil's x and y), and from where they land on the stop, in units
// of rs, build the affine map of the relative pupil coordinates
//     x_rel -> x_rel * x_gain,   y_rel -> y_rel * y_gain + y_shift            (rtl:196-206)
// The reference does this with three nested eager traces and two autograd backward calls.
// A tee ray that fails contributes a zero step, like the reference's isfinite() guard (rtl:189-190).
template <class T>
__device__ __forceinline__ bool trace_to_stop(const TlLens &ln, const float *mu_row, bool d_line, int b,
                                              int n_front, bool allow_backward, Ray<T> &r) {
  T min_cos2(1.0f);
  float min_travel = 3.0e38f;
  bool prev_live = false;
  float n_prev = 1.0f;
  for (int s = 0; s < n_front; ++s) {
    const int64_t i = (int64_t)b * ln.L + s;
    const bool live = ln.mask[i] != 0;
    float ratio = 1.0f;
    if (d_line) {                               // compute_pupil_radius traces at the d line: n = nd
      const float n_here = live ? nd_at(ln, b, s) : 1.0f;
      ratio = n_prev / n_here;
      n_prev = n_here;
    } else if (live) {
      ratio = mu_row[s];
    }
    T travel;
    fast_surface(r, T(live ? ln.c[i] : 0.f), T(ratio), T(ratio * ratio), T(live ? ln.t[i] : 0.f), min_cos2,
                 travel);
    if (s > 0 && prev_live) min_travel = fminf(min_travel, travel.v);
    prev_live = live;
  }
  const T travel = fast_image(r);
  if (n_front > 0 && prev_live) min_travel = fminf(min_travel, travel.v);
  bool ok = min_cos2.v > kGuard && fabsf(r.x.v + r.y.v) < 3.0e38f;
  if (!allow_backward && min_travel < 0.f) ok = false;
  return ok;
}

__device__ __forceinline__ void aim_one(const TlLens &ln, const float *mu, const float *z, const float *cy,
                                        const float *half_epd, const float *vig, int mode, int allow_backward,
                                        float *aim, int i) {
  const int w = i % ln.W, f = (i / ln.W) % ln.F, b = i / (ln.W * ln.F);
  const int n_front = min(ln.stop_idx[b], ln.L);
  const float h = half_epd[b], z0 = z[b];
  const float *mu_row = mu + ((int64_t)b * ln.W + w) * ln.L;
  float rs;
  if (mode == TL_AIM_PARAXIAL) {
    // stop radius = first-order magnification of the front group (the A element of its ABCD product,
    // compute_magnification rt_tf:765-777) times the pupil radius (rtl:138-140)
    Abcd m{1.f, 0.f, 0.f, 1.f};
    for (int s = 0; s < n_front; ++s) {
      const Abcd q = slot_matrix(ln, b, s, nullptr, nullptr);
      m = Abcd{q.a * m.a + q.b * m.c, q.a * m.b + q.b * m.d, q.c * m.a + q.d * m.c, q.c * m.b + q.d * m.d};
    }
    rs = m.a * h;
  } else {
    // stop radius: marginal ray of the axial field at the d line (compute_pupil_radius rtl:834-844)
    Ray<D2> m{D2(0.f), D2(h), D2(z0), D2(0.f), D2(0.f), D2(1.0f)};
    trace_to_stop<D2>(ln, mu_row, true, b, n_front, allow_backward != 0, m);
    rs = m.y.v;
  }
  // the tee rays of this (field, wavelength), vignetted like the pupil itself (rtl:154-160)
  const float dir_y = cy[(int64_t)b * ln.F + f];
  float tee_x[3] = {0.f, 0.f, 1.f}, tee_y[3] = {-1.f, 1.f, 0.f};
  if (vig) {
    const float *v = vig + ((int64_t)b * ln.F + f) * 3;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      tee_x[q] = __fmul_rn(tee_x[q], v[0]);
      tee_y[q] = __fadd_rn(__fmul_rn(tee_y[q], v[1]), v[2]);
    }
  }
  float step_x[3], step_y[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    Ray<D2> r{D2(tee_x[q] * h, h, 0.f), D2(tee_y[q] * h, 0.f, h), D2(z0), D2(0.f), D2(dir_y),
              fast_cz0(D2(0.f), D2(dir_y))};
    const bool ok = trace_to_stop<D2>(ln, mu_row, false, b, n_front, allow_backward != 0, r);
    // d(xs_rel)/d(pupil) summed over the two outputs, as the reference's two backward calls leave it
    const float xs = ok ? r.x.v / rs : 0.f, ys = ok ? r.y.v / rs : 0.f;
    const float slope_x = ok ? (r.x.a + r.y.a) / rs : 0.f, slope_y = ok ? (r.x.b + r.y.b) / rs : 0.f;
    const float sx = -(xs - tee_x[q]) / slope_x, sy = -(ys - tee_y[q]) / slope_y;
    step_x[q] = (fabsf(sx) <= 3.0e38f) ? sx : 0.f;            // not finite -> no correction
    step_y[q] = (fabsf(sy) <= 3.0e38f) ? sy : 0.f;
  }
  const float y_lo = tee_y[0], y_hi = tee_y[1], x_sag = tee_x[2];
  float *out = aim + (int64_t)i * 3;
  out[0] = (x_sag + step_x[2]) / x_sag;
  out[1] = (y_hi + step_y[1] - (y_lo + step_y[0])) / (y_hi - y_lo);
  out[2] = (y_lo * step_y[1] - y_hi * step_y[0]) / (y_lo - y_hi);
}

__global__ void k_aim(TlLens ln, const float *mu, const float *z, const float *cy, const float *half_epd,
                      const float *vig, int mode, int allow_backward, float *aim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ln.B * ln.F * ln.W) aim_one(ln, mu, z, cy, half_epd, vig, mode, allow_backward, aim, i);
}

// Staging, ray aiming and the reference heights of a lens as ONE launch (a step of the fused lens
// pass is a handful of launches of which only one is long: every launch saved is ~3 us of a ~270 us
// step).  One CTA per lens; the phases see each other's global writes through the barriers.
__global__ void __launch_bounds__(128)
k_stage_ref(TlLens ln, TlProblem pb, float *mu, float *z, float *cy, float *half_epd, float *aim,
            const float *vig, int mode, int allow_backward, float *ref_y) {
  __shared__ LensCopy copy;
  __shared__ float mu0[kStageMaxSurfaces];
  const int b = blockIdx.x;
  const TlLens lens = lens_in_shared(ln, b, copy);
  stage_fwd_lens(lens, b, mu, z, cy, half_epd, mu0);
  __threadfence_block();
  __syncthreads();
  if (aim) {
    for (int i = threadIdx.x; i < ln.F * ln.W; i += blockDim.x)
      aim_one(lens, mu, z, cy, half_epd, vig, mode, allow_backward, aim, b * ln.F * ln.W + i);
    __threadfence_block();
    __syncthreads();
  }
  if (ref_y)      // (pb.c / pb.t are the lens' own rows: the shared copy serves the per-surface loop)
    for (int f = threadIdx.x; f < pb.F; f += blockDim.x)
      chief_ray_one(pb, b * pb.F + f, ref_y, copy.f, copy.f + ln.L, mu0);
}

// Chain rule of the staging for lens b: `gmu_b` = d loss / d mu[b] ([W,L]), `gz_b` = d loss / d z[b]; ADDS
// to gc, gt, gnd, gv (global [B,L] arrays).
__device__ __forceinline__ void stage_bwd_lens(const TlLens &ln, int b, const float *gmu_b, float gz_b, float *gc,
                                               float *gt, float *gnd, float *gv) {
  // ---- through mu[w,s] = n[w,s-1] / n[w,s]: one thread per slot gathers over wavelengths
  for (int s = threadIdx.x; s < ln.L; s += blockDim.x) {
    const int64_t i = (int64_t)b * ln.L + s;
    if (!ln.mask_g[i]) continue;
    float g_nd = 0.f, g_v = 0.f;
    for (int w = 0; w < ln.W; ++w) {
      const float wl = ln.wavelengths[w];
      float dn_dnd, dn_dv;
      const float n_s = index_at(ln, b, s, wl, &dn_dnd, &dn_dv);
      const float n_in = (s == 0) ? 1.0f : index_at(ln, b, s - 1, wl, nullptr, nullptr);
      const int j = w * ln.L + s;
      float g_n = -gmu_b[j] * n_in / (n_s * n_s);                   // as denominator of mu[s]
      if (s + 1 < ln.L)                                             // as numerator of mu[s+1]
        g_n += gmu_b[j + 1] / index_at(ln, b, s + 1, wl, nullptr, nullptr);
      g_nd += g_n * dn_dnd;
      g_v += g_n * dn_dv;
    }
    gnd[i] += g_nd;
    gv[i] += g_v;
  }
  __syncthreads();
  // ---- through z = M01 / M00, M = M_{n-1} ... M_0 (thread 0, reverse sweep over the chain)
  if (threadIdx.x != 0) return;
  const int n_front = min(min(ln.stop_idx[b], ln.L), kStageMaxSurfaces);
  if (n_front <= 0) return;
  Abcd prefix[kStageMaxSurfaces];                      // P_s = M_{s-1} ... M_0
  Abcd m{1.f, 0.f, 0.f, 1.f};
  for (int s = 0; s < n_front; ++s) {
    prefix[s] = m;
    const Abcd q = slot_matrix(ln, b, s, nullptr, nullptr);
    m = Abcd{q.a * m.a + q.b * m.c, q.a * m.b + q.b * m.d, q.c * m.a + q.d * m.c,
             q.c * m.b + q.d * m.d};
  }
  const float gzz = gz_b;
  Abcd g{-gzz * m.b / (m.a * m.a), gzz / m.a, 0.f, 0.f};        // adjoint of the full product
  for (int s = n_front - 1; s >= 0; --s) {
    // total = Q_s M_s P_s ; g holds Q_s^T G ; d M_s = g P_s^T
    const Abcd p = prefix[s];
    const Abcd dm{g.a * p.a + g.b * p.b, g.a * p.c + g.b * p.d, g.c * p.a + g.d * p.b,
                  g.c * p.c + g.d * p.d};
    float ratio, power;
    const Abcd q = slot_matrix(ln, b, s, &ratio, &power);
    const int64_t i = (int64_t)b * ln.L + s;
    if (ln.mask[i]) {
      const float tt = ln.t[i], cc = ln.c[i];
      gt[i] += dm.a * power + dm.b * ratio;
      const float g_power = dm.a * tt + dm.c;
      float g_ratio = dm.b * tt + dm.d + g_power * cc;
      gc[i] += g_power * (ratio - 1.0f);
      const float n_out = nd_at(ln, b, s);
      if (ln.mask_g[i]) gnd[i] += -g_ratio * ratio / n_out;
      if (s > 0 && ln.mask_g[i - 1]) gnd[i - 1] += g_ratio / n_out;
    }
    g = Abcd{q.a * g.a + q.c * g.c, q.a * g.b + q.c * g.d, q.b * g.a + q.d * g.c,
             q.b * g.b + q.d * g.d};                             // M_s^T g
  }
}

__global__ void k_stage_bwd(TlLens ln, const float *gmu, const float *gz, float *gc, float *gt,
                            float *gnd, float *gv) {
  const int b = blockIdx.x;
  stage_bwd_lens(ln, b, gmu + (int64_t)b * ln.W * ln.L, gz[b], gc, gt, gnd, gv);
}

// Finalize + staging chain rule of a lens as ONE launch: moments -> rms, rms_field and the gradients
// of rms[b] w.r.t. the padded lens tensors c, t, nd, v [B,L] (what k_spot_finalize, two fills and
// k_stage_bwd did in four).  out.gmu / out.gz are scratch here.
template <bool STAGED>
__global__ void k_lens_finalize(const double *mom_global, const float *ref_y, int B, int F, int W, int S,
                                double n_rays, TlSpotOut out, TlLens ln, float *gnd, float *gv) {
  spot_finalize_lens<STAGED>(mom_global, ref_y, B, F, W, S, n_rays, 1, out);
  const int b = blockIdx.x;
  // The chain rule's reverse sweep over the surfaces in front of the stop is serial (thread 0) and adds
  // into gc / gt / gnd as it goes: on global memory every `+=` is a ~0.7 us round trip of a dependent
  // chain (12 us of a 240 us step).  It runs on a shared-memory copy of the four gradient rows instead.
  __shared__ float rows4[4][kStageMaxSurfaces];
  __shared__ LensCopy copy;
  __threadfence_block();
  __syncthreads();
  const int64_t bias = (int64_t)b * ln.L;
  for (int s_ = threadIdx.x; s_ < ln.L; s_ += blockDim.x) {
    rows4[0][s_] = out.gc[bias + s_];
    rows4[1][s_] = out.gt[bias + s_];
    rows4[2][s_] = 0.f;
    rows4[3][s_] = 0.f;
  }
  const TlLens lens = lens_in_shared(ln, b, copy);        // (ends with the barrier the phases need)
  stage_bwd_lens(lens, b, out.gmu + (int64_t)b * ln.W * ln.L, out.gz[b], rows4[0] - bias, rows4[1] - bias,
                 rows4[2] - bias, rows4[3] - bias);
  __syncthreads();
  for (int s_ = threadIdx.x; s_ < ln.L; s_ += blockDim.x) {
    out.gc[bias + s_] = rows4[0][s_];
    out.gt[bias + s_] = rows4[1][s_];
    gnd[bias + s_] = rows4[2][s_];
    gv[bias + s_] = rows4[3][s_];
  }
}

// --------------------------------------------------------------------------
// host-side planning
// --------------------------------------------------------------------------
typedef void (*AdjKernelPtr)(TlProblem, AdjArgs);

bool is_general(const TlProblem &pb) { return pb.k || pb.a || pb.sd; }

int validate(const TlProblem *pb, int max_s) {
  if (!pb) return fail(TL_ERR_INVALID, "problem is NULL%s");
  if (pb->B < 1 || pb->F < 1 || pb->P < 1 || pb->W < 1 || pb->S < 1)
    return fail(TL_ERR_INVALID, "B, F, P, W, S must all be >= 1%s");
  if (pb->S > (is_general(*pb) ? TL_MAX_SURFACES_GEN : max_s))
    return fail(TL_ERR_INVALID, "too many surfaces for this entry point%s");
  if (!pb->x.ptr || !pb->y.ptr || !pb->z.ptr || !pb->cx.ptr || !pb->cy.ptr || !pb->c || !pb->t ||
      !pb->mu || !pb->live)
    return fail(TL_ERR_INVALID, "NULL input pointer%s");
  if (pb->arith != TL_ARITH_GUARDED && pb->arith != TL_ARITH_EXACT)
    return fail(TL_ERR_INVALID, "unknown arithmetic policy%s");
  if ((int64_t)pb->B * pb->F * pb->W > (1 << 24))
    return fail(TL_ERR_INVALID, "B*F*W too large%s");
  if ((pb->aim || pb->vig) && is_general(*pb))
    return fail(TL_ERR_INVALID, "the ray-aiming / vignetting maps are for spherical lenses only (not with k / a / sd)%s");
  return TL_OK;
}

// K1 grid: chunks per (b,f,w) so that the machine is filled a few times over
int n_acc_gen(int S, int want_grad) { return want_grad ? kGenSlots * S + 5 : 3; }

// plan of the general-surface spot kernel
struct GenPlan {
  AdjKernelPtr kernel = nullptr;
  int lanes = 2, n_blocks = 1, groups_per_row = 1, max_seg = 1, n_acc = 0;
  size_t smem = 0, partial_bytes = 0;
};

struct FwdPlan {
  int nchunks = 1, chunk_len = 1, n_blocks = 1;
};

FwdPlan make_fwd_plan(int sms, int bfw, int n_pupil, int rays_per_pass, int ctas_per_sm) {
  FwdPlan pl;
  const int64_t want_blocks = (int64_t)sms * ctas_per_sm * 4;
  int64_t nchunks = (want_blocks + bfw - 1) / bfw;
  const int64_t max_chunks = ((int64_t)n_pupil + rays_per_pass - 1) / rays_per_pass;
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

// ---- K2/K3 variants -------------------------------------------------------
typedef AdjKernelPtr AdjKernel;

struct AdjVariant {
  AdjKernel kernel = nullptr;
  int lanes = 2;          // rays per thread and group
  int state_lanes = 2;    // lanes' worth of parked state per thread
};

// Rays per thread.  Four (two interleaved pairs) hides the latency of the
// dependent FMA/MUFU chain best when the accumulators still fit the register file
// (S <= 12 for the fused pass); larger surface counts use two.  TL_LANES overrides
// the choice for experiments.
int pick_lanes(int mode, int S) {
  int lanes = 2;
  if (mode == MODE_SPOT_EVAL) lanes = 4;
  else if (mode == MODE_SPOT_GRAD) lanes = (S <= 12) ? TL_DEFAULT_LANES_SPOT : 2;
  else lanes = (S <= 12) ? TL_DEFAULT_LANES_BWD : 2;
  if (const char *env = getenv("TL_LANES")) {
    const int v = atoi(env);
    if ((v == 2 || v == 4) && mode != MODE_SPOT_EVAL && !(mode == MODE_BWD && S > 16)) lanes = v;
  }
  return lanes;
}

template <int MODE, class V>
AdjKernel adj_kernel_for(int S) {
  if constexpr (MODE == MODE_SPOT_EVAL) {
    return k_trace_adj<1, MODE_SPOT_EVAL, V>;
  } else {
    if (S <= 4) return k_trace_adj<4, MODE, V>;
    if (S <= 8) return k_trace_adj<8, MODE, V>;
    if (S <= 12) return k_trace_adj<12, MODE, V>;
    return k_trace_adj<16, MODE, V>;
  }
}

// MODE_BWD with seeds on the aggregate=True stacks: two rays per thread
template <int PEN>
AdjKernel pen_kernel_for(int S) {
  if (S <= 4) return k_trace_adj<4, MODE_BWD, f2, PEN>;
  if (S <= 8) return k_trace_adj<8, MODE_BWD, f2, PEN>;
  if (S <= 12) return k_trace_adj<12, MODE_BWD, f2, PEN>;
  if (S <= 16) return k_trace_adj<16, MODE_BWD, f2, PEN>;
  return k_trace_adj<32, MODE_BWD, f2, PEN>;
}

AdjVariant select_adj(int mode, int S, int pen = PEN_NONE) {
  AdjVariant v;
  if (pen != PEN_NONE && mode == MODE_BWD) {
    v.lanes = v.state_lanes = 2;
    v.kernel = pen == PEN_SUM ? pen_kernel_for<PEN_SUM>(S) : pen_kernel_for<PEN_SEEDED>(S);
    return v;
  }
  v.lanes = pick_lanes(mode, S);
  if (mode == MODE_SPOT_EVAL) {
    v.kernel = adj_kernel_for<MODE_SPOT_EVAL, f4>(S);
  } else if (mode == MODE_SPOT_GRAD) {
    v.kernel = v.lanes == 4 ? adj_kernel_for<MODE_SPOT_GRAD, f4>(S) : adj_kernel_for<MODE_SPOT_GRAD, f2>(S);
  } else if (S > 16) {
    v.kernel = k_trace_adj<32, MODE_BWD, f2>;
  } else {
    v.kernel = v.lanes == 4 ? adj_kernel_for<MODE_BWD, f4>(S) : adj_kernel_for<MODE_BWD, f2>(S);
  }
  v.state_lanes = v.lanes;
  return v;
}

struct AdjPlan {
  AdjVariant variant;
  int n_blocks = 1, groups_per_row = 1, max_seg = 1, n_acc = 0;
  size_t smem = 0;
  size_t partial_bytes = 0;
};

size_t align8(size_t v) { return (v + 7) & ~(size_t)7; }

int plan_adj(const TlProblem &pb, int mode, AdjPlan &pl, int pen = PEN_NONE) {
  DeviceInfo info;
  int rc = device_info(info);
  if (rc) return rc;
  pl.variant = select_adj(mode, pb.S, pen);
  pl.n_acc = n_acc_of(mode, pb.S) + (pen == PEN_SUM ? 1 : 0);
  const int lanes = pl.variant.lanes;
  const size_t table = ((5 * (size_t)pb.S + 3) & ~(size_t)3) * sizeof(float);
  const size_t state = mode != MODE_SPOT_EVAL
                           ? (size_t)4 * pb.S * kTraceThreads * pl.variant.state_lanes * sizeof(float) : 0;
  const size_t red = (size_t)(kTraceThreads / 32) * pl.n_acc * sizeof(float) + 16;
  pl.smem = table + (state > red ? state : red);
  if (pl.smem > 227 * 1024) return fail(TL_ERR_INVALID, "surface count needs too much shared memory%s");
  if (pl.smem > 48 * 1024)
    TL_CHECK_CUDA(cudaFuncSetAttribute((const void *)pl.variant.kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  int per_sm = 0;
  TL_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)pl.variant.kernel,
                                                              kTraceThreads, pl.smem));
  if (per_sm < 1) return fail(TL_ERR_CUDA, "kernel does not fit on an SM%s");
  const int n_pupil = pb.p_end - pb.p_begin;
  const int group = kTraceThreads * lanes;
  pl.groups_per_row = (n_pupil + group - 1) / group;
  const int64_t total = (int64_t)pb.B * pb.F * pb.W * pl.groups_per_row;
  int64_t n_blocks = (int64_t)info.sms * per_sm;
  if (n_blocks > total) n_blocks = total;
  pl.n_blocks = (int)n_blocks;
  const int64_t len = (total + n_blocks - 1) / n_blocks;          // longest slice
  pl.max_seg = (int)((len - 1 + pl.groups_per_row - 1) / pl.groups_per_row + 1);
  pl.partial_bytes = align8((size_t)pl.n_blocks * pl.max_seg * pl.n_acc * sizeof(double));
  return TL_OK;
}

int launch_adj(const TlProblem &pb, AdjArgs &args, const AdjPlan &pl, cudaStream_t stream) {
  args.groups_per_row = pl.groups_per_row;
  args.max_seg = pl.max_seg;
  args.n_acc = pl.n_acc;
  TlProblem pb_copy = pb;
  void *params[] = {(void *)&pb_copy, (void *)&args};
  TL_CHECK_CUDA(cudaLaunchKernel((const void *)pl.variant.kernel, dim3(pl.n_blocks), dim3(kTraceThreads),
                                 params, pl.smem, stream));
  g_launches++;
  return TL_OK;
}

int reduce_rows(const AdjPlan &pl, const double *partial, double *rows_out, int rows,
                cudaStream_t stream) {
  const int64_t n = (int64_t)rows * pl.n_acc;
  k_reduce_rows<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(partial, rows_out, rows,
                                                                 pl.groups_per_row, pl.n_blocks,
                                                                 pl.max_seg, pl.n_acc);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

#include "spot_rev.cuh"
#include "psf_kernels.cuh"
#include "paraxial.cuh"

// K3c: the backward kernels (MODE_BWD of k_trace_adj: plain, with seeded stacks, and the fused
// penalty pass PEN_SUM) with a warp per row, for the same many-short-rows workload as k_spot_rows.
// Row layout [3S+1 (+1)]: c, t, mu per surface, z (, penalty) -- what k_bwd_finalize /
// k_penalty_finalize read.
template <int NS_MAX, class V, int PEN>
__global__ void __launch_bounds__(kTraceThreads)
k_bwd_rows(TlProblem pb, AdjArgs args, double *rows, int n_acc) {
  extern __shared__ float smem[];
  constexpr int N = LaneCount<V>::value;
  constexpr int kWarps = kTraceThreads / 32;
  const int S = pb.S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool allow_backward = pb.allow_backward_rays != 0;
  const size_t tab_floats = (table_floats(S) + 3) & ~(size_t)3;
  const size_t state_floats = (size_t)4 * S * 32 * N;
  float *wbase = smem + warp * (tab_floats + state_floats);
  V *state = reinterpret_cast<V *>(wbase + tab_floats) + lane;
  constexpr int stride = 32;
  const int n_rows = pb.B * pb.F * pb.W;
  const int groups = (pb.p_end - pb.p_begin + 32 * N - 1) / (32 * N);
  const int64_t plane = (int64_t)pb.B * pb.F * pb.P * pb.W;

  for (int row = blockIdx.x * kWarps + warp; row < n_rows; row += gridDim.x * kWarps) {
    const int w = row % pb.W;
    const int f = (row / pb.W) % pb.F;
    const int b = row / (pb.W * pb.F);
    const Table tab = load_table_warp(wbase, pb, b, w, lane);
    const float xy_scale = pb.xy_scale ? pb.xy_scale[b] : 1.0f;
    float acc_c[NS_MAX], acc_t[NS_MAX], acc_mu[NS_MAX];
#pragma unroll
    for (int k = 0; k < NS_MAX; ++k) acc_c[k] = acc_t[k] = acc_mu[k] = 0.f;
    float acc_z = 0.f, m_pen = 0.f;

    for (int j = 0; j < groups; ++j) {
      const int p_base = pb.p_begin + j * (32 * N) + lane;
      if (p_base >= pb.p_end) continue;
      bool has[N];
      int64_t o[N];
      V x, y, z, cx, cy;
#pragma unroll
      for (int l = 0; l < N; ++l) {
        const int p = p_base + l * 32;
        has[l] = p < pb.p_end;
        const int q = has[l] ? p : p_base;
        o[l] = (((int64_t)b * pb.F + f) * pb.P + q) * pb.W + w;
        float px, py;
        load_pupil_point(pb, b, f, q, w, xy_scale, px, py);
        lane_set(x, l, px);
        lane_set(y, l, py);
        lane_set(z, l, pb.z.ptr[offset_of(pb.z, b, f, q, w)]);
        lane_set(cx, l, pb.cx.ptr[offset_of(pb.cx, b, f, q, w)]);
        lane_set(cy, l, pb.cy.ptr[offset_of(pb.cy, b, f, q, w)]);
      }
      TracedN<V> tr = trace_guarded<true, V, PEN != PEN_NONE>(x, y, z, cx, cy, tab, S, allow_backward, pb.arith,
                                                            state, stride);
      bool live[N], any_live = false, all_ok = true;
      unsigned bits[N], flips[N];
#pragma unroll
      for (int l = 0; l < N; ++l) {
        live[l] = tr.ok[l] && has[l];
        any_live = any_live || live[l];
        all_ok = all_ok && tr.ok[l];
        bits[l] = (PEN != PEN_NONE && has[l]) ? tr.ok_bits[l][0] : 0u;
        flips[l] = PEN != PEN_NONE ? tr.ok_bits[l][1] : 0u;
      }
      Ray<V> a{V(0.f), V(0.f), V(0.f), V(0.f), V(0.f), V(0.f)};
      if (PEN != PEN_NONE || any_live) {
        if (PEN == PEN_NONE && !all_ok) mirror_live_lane<V>(state, stride, S, tr.ok, tr.pre, z, tr.x, tr.y);
        V sx(0.f), sy(0.f), scx(0.f), scy(0.f);
        if (PEN != PEN_SUM) {
#pragma unroll
          for (int l = 0; l < N; ++l) {
            if (!live[l]) continue;
            if (args.seeds.gx) lane_set(sx, l, args.seeds.gx[o[l]]);
            if (args.seeds.gy) lane_set(sy, l, args.seeds.gy[o[l]]);
            if (args.seeds.gcx) lane_set(scx, l, args.seeds.gcx[o[l]]);
            if (args.seeds.gcy) lane_set(scy, l, args.seeds.gcy[o[l]]);
          }
        }
        Sweep<V> sw = sweep_begin(tr.pre, tr.x, tr.y, sx, sy, scx, scy);
#pragma unroll
        for (int k = NS_MAX - 1; k >= 0; --k) {
          if (k >= S) continue;
          const V *slot = state + (size_t)k * 4 * stride;
          SurfaceGrad<V> g;
          if constexpr (PEN != PEN_NONE) {
            V pz(0.f), pth(0.f), pthp(0.f), branch(1.0f);
            float dead_gt[N];
#pragma unroll
            for (int l = 0; l < N; ++l) {
              dead_gt[l] = 0.f;
              if ((flips[l] >> k) & 1u) lane_set(branch, l, -1.0f);
              if (!has[l]) continue;
              const int64_t at = (int64_t)k * plane + o[l];
              if (PEN == PEN_SUM) {
                if ((bits[l] >> k) & 1u) {
                  lane_set(pz, l, 1.0f);
                  lane_set(pth, l, 1.0f);
                  lane_set(pthp, l, 1.0f);
                } else {      // failed ray: theta = theta' = 1, z = 0 - t[k] (rtl:639, :653-654)
                  const float z_dead = -tab.t[k];
                  m_pen += 2.0f + fmaxf(z_dead, 0.0f);
                  if (z_dead > 0.f) dead_gt[l] = -1.0f;
                }
              } else if ((bits[l] >> k) & 1u) {
                if (args.seeds.gz_relu) lane_set(pz, l, args.seeds.gz_relu[at]);
                if (args.seeds.gtheta) lane_set(pth, l, args.seeds.gtheta[at]);
                if (args.seeds.gtheta_prime) lane_set(pthp, l, args.seeds.gtheta_prime[at]);
              } else if (args.seeds.gz_relu && -tab.t[k] > 0.f) {
                dead_gt[l] = -args.seeds.gz_relu[at];
              }
            }
            V cos_in, cos_out, z_behind;
            g = sweep_sphere_pen(sw, slot[0], slot[stride], slot[2 * stride], slot[3 * stride], V(tab.c[k]),
                                 V(tab.t[k]), V(tab.mu[k]), V(tab.mu2[k]), pz, pth, pthp, branch, cos_in, cos_out,
                                 z_behind);
#pragma unroll
            for (int l = 0; l < N; ++l) {
              if ((bits[l] >> k) & 1u) {
                if (PEN == PEN_SUM)
                  m_pen += (fast_angle_norm(lane_get(cos_in, l)) + fast_angle_norm(lane_get(cos_out, l))) +
                           fmaxf(lane_get(z_behind, l), 0.0f);
                continue;
              }
              lane_set(g.c, l, 0.f);
              lane_set(g.mu, l, 0.f);
              lane_set(g.t, l, dead_gt[l]);
              lane_set(sw.gr.x, l, 0.f); lane_set(sw.gr.y, l, 0.f); lane_set(sw.gr.z, l, 0.f);
              lane_set(sw.gd.x, l, 0.f); lane_set(sw.gd.y, l, 0.f); lane_set(sw.gd.z, l, 0.f);
            }
          } else {
            g = sweep_sphere(sw, slot[0], slot[stride], slot[2 * stride], slot[3 * stride], V(tab.c[k]),
                             V(tab.t[k]), V(tab.mu[k]), V(tab.mu2[k]));
          }
          acc_c[k] += lane_sum(g.c);
          acc_t[k] += lane_sum(g.t);
          acc_mu[k] += lane_sum(g.mu);
        }
        sweep_end(sw, z, a.x, a.y, a.z, a.cx, a.cy);
        acc_z += lane_sum(a.z);
      }
      if (PEN != PEN_SUM) {
#pragma unroll
        for (int l = 0; l < N; ++l) {
          if (!has[l]) continue;
          if (args.grads.gx) args.grads.gx[o[l]] = lane_get(a.x, l) * xy_scale;
          if (args.grads.gy) args.grads.gy[o[l]] = lane_get(a.y, l) * xy_scale;
          if (args.grads.gz) args.grads.gz[o[l]] = lane_get(a.z, l);
          if (args.grads.gcx) args.grads.gcx[o[l]] = lane_get(a.cx, l);
          if (args.grads.gcy) args.grads.gcy[o[l]] = lane_get(a.cy, l);
        }
      }
    }

    // three values per surface, ten surfaces per transpose: lane 3 q + typ -> slot typ * S + k
    double *dst = rows + (int64_t)row * n_acc;
    constexpr int kPerBatch = 10;
#pragma unroll
    for (int i = 0; i < (NS_MAX + kPerBatch - 1) / kPerBatch; ++i) {
      float v[32];
#pragma unroll
      for (int q = 0; q < kPerBatch; ++q) {
        const int k = i * kPerBatch + q;
        const bool in = k < NS_MAX;
        v[3 * q + 0] = in ? acc_c[k < NS_MAX ? k : 0] : 0.f;
        v[3 * q + 1] = in ? acc_t[k < NS_MAX ? k : 0] : 0.f;
        v[3 * q + 2] = in ? acc_mu[k < NS_MAX ? k : 0] : 0.f;
      }
      v[30] = v[31] = 0.f;
      const float total = warp_transpose_sum(v, lane);
      const int k = i * kPerBatch + lane / 3, typ = lane % 3;
      if (lane < 30 && k < S) dst[typ * S + k] = (double)total;
    }
    const float t0 = warp_sum(acc_z), t1 = warp_sum(m_pen);
    if (lane == 0) {
      dst[3 * S] = (double)t0;
      if (PEN == PEN_SUM) dst[3 * S + 1] = (double)t1;
    }
  }
}

// ---- K3b dispatch: warp-per-row spot pass for short pupil slices -------------------------
typedef void (*RowsKernelPtr)(TlProblem, const float *, double *, int);
constexpr int kLegacySpotSurfaces = 16; // k_spot_rows / k_trace_adj<SPOT_GRAD> keep 6 S sums in registers: up to 16
                                        // surfaces; beyond (to TL_MAX_SURFACES_SPOT) only k_spot_rev's TMEM variants
constexpr int kRowsMaxPupil = 1024;     // measured crossover (S = 7, 1.5 M rays): rows 112 vs CTA 92 G events/s at
                                        // P = 1024, 77 vs 102 at P = 2025 (tools/rows_crossover.sh)

RowsKernelPtr rows_kernel_for(int S, int want_grad) {
  if (!want_grad) return k_spot_rows<1, false, f2>;
  if (S <= 4) return k_spot_rows<4, true, f2>;
  if (S <= 8) return k_spot_rows<8, true, f2>;
  if (S <= 12) return k_spot_rows<12, true, f2>;
  return k_spot_rows<16, true, f2>;
}

bool use_rows_kernel(const TlProblem &pb, int want_grad) {
  if (getenv("TL_NO_ROWS")) return false;
  const int n_pupil = pb.p_end - pb.p_begin;
  const int64_t rows = (int64_t)pb.B * pb.F * pb.W;
  if (want_grad && pb.S > kLegacySpotSurfaces) return false;      // (instantiated up to 16 surfaces)
  int max_pupil = kRowsMaxPupil;
  if (const char *env = getenv("TL_ROWS_MAX_PUPIL")) max_pupil = atoi(env);     // crossover experiments
  return n_pupil <= max_pupil && rows >= 64;
}

int launch_spot_rows(const TlProblem &pb, int want_grad, const float *ref_y, double *moments,
                     cudaStream_t stream) {
  DeviceInfo info;
  int rc = device_info(info);
  if (rc) return rc;
  RowsKernelPtr kernel = rows_kernel_for(pb.S, want_grad);
  const size_t tab_floats = (5 * (size_t)pb.S + 3) & ~(size_t)3;
  const size_t state_floats = want_grad ? (size_t)4 * pb.S * 32 * 2 : 0;
  const size_t smem = (kTraceThreads / 32) * (tab_floats + state_floats) * sizeof(float);
  if (smem > 227 * 1024) return fail(TL_ERR_INVALID, "surface count needs too much shared memory%s");
  if (smem > 48 * 1024)
    TL_CHECK_CUDA(cudaFuncSetAttribute((const void *)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
  int per_sm = 0;
  TL_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)kernel, kTraceThreads,
                                                              smem));
  if (per_sm < 1) return fail(TL_ERR_CUDA, "kernel does not fit on an SM%s");
  const int64_t rows = (int64_t)pb.B * pb.F * pb.W;
  int64_t n_blocks = (int64_t)info.sms * per_sm;
  const int64_t need = (rows + kTraceThreads / 32 - 1) / (kTraceThreads / 32);
  if (n_blocks > need) n_blocks = need;
  const int n_acc = n_acc_of(want_grad ? MODE_SPOT_GRAD : MODE_SPOT_EVAL, pb.S);
  TlProblem pb_copy = pb;
  void *params[] = {(void *)&pb_copy, (void *)&ref_y, (void *)&moments, (void *)&n_acc};
  TL_CHECK_CUDA(cudaLaunchKernel((const void *)kernel, dim3((unsigned)n_blocks), dim3(kTraceThreads), params,
                                 smem, stream));
  g_launches++;
  return TL_OK;
}

typedef void (*BwdRowsKernelPtr)(TlProblem, AdjArgs, double *, int);

template <int PEN>
BwdRowsKernelPtr bwd_rows_kernel_for(int S) {
  if (S <= 4) return k_bwd_rows<4, f2, PEN>;
  if (S <= 8) return k_bwd_rows<8, f2, PEN>;
  if (S <= 12) return k_bwd_rows<12, f2, PEN>;
  return k_bwd_rows<16, f2, PEN>;
}

// warp-per-row backward (S <= 16): rows[B*F*W][3S+1 (+1 for PEN_SUM)]
int launch_bwd_rows(const TlProblem &pb, const AdjArgs &args, int pen, double *rows, cudaStream_t stream) {
  DeviceInfo info;
  int rc = device_info(info);
  if (rc) return rc;
  BwdRowsKernelPtr kernel = pen == PEN_SUM ? bwd_rows_kernel_for<PEN_SUM>(pb.S)
                            : pen == PEN_SEEDED ? bwd_rows_kernel_for<PEN_SEEDED>(pb.S)
                                                : bwd_rows_kernel_for<PEN_NONE>(pb.S);
  const size_t tab_floats = (5 * (size_t)pb.S + 3) & ~(size_t)3;
  const size_t smem = (kTraceThreads / 32) * (tab_floats + (size_t)4 * pb.S * 32 * 2) * sizeof(float);
  if (smem > 48 * 1024)
    TL_CHECK_CUDA(cudaFuncSetAttribute((const void *)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
  int per_sm = 0;
  TL_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)kernel, kTraceThreads,
                                                              smem));
  if (per_sm < 1) return fail(TL_ERR_CUDA, "kernel does not fit on an SM%s");
  const int64_t n_rows = (int64_t)pb.B * pb.F * pb.W;
  int64_t n_blocks = (int64_t)info.sms * per_sm;
  const int64_t need = (n_rows + kTraceThreads / 32 - 1) / (kTraceThreads / 32);
  if (n_blocks > need) n_blocks = need;
  const int n_acc = 3 * pb.S + 1 + (pen == PEN_SUM ? 1 : 0);
  TlProblem pb_copy = pb;
  AdjArgs args_copy = args;
  void *params[] = {(void *)&pb_copy, (void *)&args_copy, (void *)&rows, (void *)&n_acc};
  TL_CHECK_CUDA(cudaLaunchKernel((const void *)kernel, dim3((unsigned)n_blocks), dim3(kTraceThreads), params,
                                 smem, stream));
  g_launches++;
  return TL_OK;
}

int plan_gen(const TlProblem &pb, int want_grad, GenPlan &pl, bool seeded = false) {
  DeviceInfo info;
  int rc = device_info(info);
  if (rc) return rc;
  // Rays per thread.  The forward-only sweep takes four; so does the fused pass with gradients when two CTAs of
  // it still fit an SM's shared memory (parked state 8 KB per surface and CTA: up to 12 surfaces) -- four rays
  // share one 24-slot butterfly and one accumulator update per adjoint step: 0.7313 -> 0.6832 ms on the config-3
  // lens, at 221 registers and 8 warps per SM against 125 registers and 16 warps for two rays per thread.
  // TL_GEN_LANES=2 forces the two-ray variant (A/B runs, tests).
  const size_t table = ((gen_table_floats(pb.S) + 3) & ~(size_t)3) * sizeof(float);
  const size_t rows = want_grad ? (size_t)(kTraceThreads / 32) * pb.S * kGenRow * sizeof(float) : 0;
  const size_t state4 = (size_t)4 * pb.S * kTraceThreads * 4 * sizeof(float);
  const char *env_lanes = getenv("TL_GEN_LANES");
  const bool four = want_grad && !seeded && !(env_lanes && atoi(env_lanes) == 2) &&
                    2 * (table + rows + state4 + 16 + 1024) <= (size_t)227 * 1024;
  pl.lanes = (!want_grad || four) ? 4 : 2;
  pl.kernel = seeded ? (AdjKernelPtr)k_trace_gen<MODE_BWD, f2>
              : !want_grad ? (AdjKernelPtr)k_trace_gen<MODE_SPOT_EVAL, f4>
              : four ? (AdjKernelPtr)k_trace_gen<MODE_SPOT_GRAD, f4>
                     : (AdjKernelPtr)k_trace_gen<MODE_SPOT_GRAD, f2>;
  pl.n_acc = n_acc_gen(pb.S, want_grad);
  const size_t state = want_grad ? (size_t)4 * pb.S * kTraceThreads * pl.lanes * sizeof(float) : 0;
  pl.smem = table + rows + state + 16;
  if (pl.smem > 227 * 1024) return fail(TL_ERR_INVALID, "surface count needs too much shared memory%s");
  if (pl.smem > 48 * 1024)
    TL_CHECK_CUDA(cudaFuncSetAttribute((const void *)pl.kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  int per_sm = 0;
  TL_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)pl.kernel,
                                                              kTraceThreads, pl.smem));
  if (per_sm < 1) return fail(TL_ERR_CUDA, "kernel does not fit on an SM%s");
  const int group = kTraceThreads * pl.lanes;
  pl.groups_per_row = (pb.p_end - pb.p_begin + group - 1) / group;
  const int64_t total = (int64_t)pb.B * pb.F * pb.W * pl.groups_per_row;
  int64_t n_blocks = (int64_t)info.sms * per_sm;
  if (n_blocks > total) n_blocks = total;
  pl.n_blocks = (int)n_blocks;
  const int64_t len = (total + n_blocks - 1) / n_blocks;
  pl.max_seg = (int)((len - 1 + pl.groups_per_row - 1) / pl.groups_per_row + 1);
  pl.partial_bytes = align8((size_t)pl.n_blocks * pl.max_seg * pl.n_acc * sizeof(double));
  return TL_OK;
}

}  // namespace

// --------------------------------------------------------------------------
// C ABI
// --------------------------------------------------------------------------
extern "C" {

int tl_abi_version(void) { return TL_ABI_VERSION; }
const char *tl_last_error(void) { return g_error; }
int64_t tl_launch_count(void) { return (int64_t)g_launches.load(); }

const char *tl_abi_describe(int32_t which) {
  static thread_local char text[1024];
  int n = 0;
  text[0] = 0;
#define TL_SIZE(T) n += snprintf(text + n, sizeof(text) - n, #T ":%zu", sizeof(T))
#define TL_OFF(T, f) n += snprintf(text + n, sizeof(text) - n, ";" #f "@%zu", offsetof(T, f))
  switch (which) {
    case 0: TL_SIZE(TlStrided); TL_OFF(TlStrided, ptr); TL_OFF(TlStrided, stride); break;
    case 1: TL_SIZE(TlProblem); TL_OFF(TlProblem, x); TL_OFF(TlProblem, y); TL_OFF(TlProblem, z); TL_OFF(TlProblem, cx); TL_OFF(TlProblem, cy); TL_OFF(TlProblem, c); TL_OFF(TlProblem, t); TL_OFF(TlProblem, mu); TL_OFF(TlProblem, live); TL_OFF(TlProblem, B); TL_OFF(TlProblem, F); TL_OFF(TlProblem, P); TL_OFF(TlProblem, W); TL_OFF(TlProblem, S); TL_OFF(TlProblem, allow_backward_rays); TL_OFF(TlProblem, arith); TL_OFF(TlProblem, p_begin); TL_OFF(TlProblem, p_end); TL_OFF(TlProblem, xy_scale); TL_OFF(TlProblem, k); TL_OFF(TlProblem, a); TL_OFF(TlProblem, sd); TL_OFF(TlProblem, aim); TL_OFF(TlProblem, vig); break;
    case 2: TL_SIZE(TlTraceOut); TL_OFF(TlTraceOut, x); TL_OFF(TlTraceOut, y); TL_OFF(TlTraceOut, cx); TL_OFF(TlTraceOut, cy); TL_OFF(TlTraceOut, ok); TL_OFF(TlTraceOut, backward); TL_OFF(TlTraceOut, opl); TL_OFF(TlTraceOut, z_relu); TL_OFF(TlTraceOut, theta); TL_OFF(TlTraceOut, theta_prime); break;
    case 3: TL_SIZE(TlSeeds); TL_OFF(TlSeeds, gx); TL_OFF(TlSeeds, gy); TL_OFF(TlSeeds, gcx); TL_OFF(TlSeeds, gcy); TL_OFF(TlSeeds, gz_relu); TL_OFF(TlSeeds, gtheta); TL_OFF(TlSeeds, gtheta_prime); TL_OFF(TlSeeds, gopl); break;
    case 4: TL_SIZE(TlGrads); TL_OFF(TlGrads, gc); TL_OFF(TlGrads, gt); TL_OFF(TlGrads, gmu); TL_OFF(TlGrads, gz_sum); TL_OFF(TlGrads, gx); TL_OFF(TlGrads, gy); TL_OFF(TlGrads, gz); TL_OFF(TlGrads, gcx); TL_OFF(TlGrads, gcy); TL_OFF(TlGrads, gk); TL_OFF(TlGrads, ga); break;
    case 5: TL_SIZE(TlSpotOut); TL_OFF(TlSpotOut, rms); TL_OFF(TlSpotOut, rms_field); TL_OFF(TlSpotOut, gc); TL_OFF(TlSpotOut, gt); TL_OFF(TlSpotOut, gmu); TL_OFF(TlSpotOut, gz); TL_OFF(TlSpotOut, gk); TL_OFF(TlSpotOut, ga); break;
    case 6: TL_SIZE(TlPenaltyOut); TL_OFF(TlPenaltyOut, penalty); TL_OFF(TlPenaltyOut, gc); TL_OFF(TlPenaltyOut, gt); TL_OFF(TlPenaltyOut, gmu); TL_OFF(TlPenaltyOut, gz); break;
    case 7: TL_SIZE(TlLens); TL_OFF(TlLens, c); TL_OFF(TlLens, t); TL_OFF(TlLens, nd); TL_OFF(TlLens, v); TL_OFF(TlLens, mask); TL_OFF(TlLens, mask_g); TL_OFF(TlLens, stop_idx); TL_OFF(TlLens, hfov); TL_OFF(TlLens, epd); TL_OFF(TlLens, rel_fields); TL_OFF(TlLens, wavelengths); TL_OFF(TlLens, B); TL_OFF(TlLens, L); TL_OFF(TlLens, F); TL_OFF(TlLens, W); break;
    case 8: TL_SIZE(TlPsf); TL_OFF(TlPsf, x); TL_OFF(TlPsf, y); TL_OFF(TlPsf, y_target); TL_OFF(TlPsf, x_incr); TL_OFF(TlPsf, y_incr); TL_OFF(TlPsf, x_size); TL_OFF(TlPsf, y_size); TL_OFF(TlPsf, G); TL_OFF(TlPsf, C); TL_OFF(TlPsf, R); TL_OFF(TlPsf, n_x_bins); TL_OFF(TlPsf, n_y_bins); break;
    case 9: TL_SIZE(TlParaxial); TL_OFF(TlParaxial, c); TL_OFF(TlParaxial, t); TL_OFF(TlParaxial, n); TL_OFF(TlParaxial, live); TL_OFF(TlParaxial, glass); TL_OFF(TlParaxial, B); TL_OFF(TlParaxial, L); TL_OFF(TlParaxial, mode); break;
    default: return nullptr;
  }
#undef TL_SIZE
#undef TL_OFF
  return text;
}

int tl_trace_fwd(const TlProblem *pb, const TlTraceOut *out, void *stream_) {
  int rc = validate(pb, TL_MAX_SURFACES_FWD);
  if (rc) return rc;
  if (!out || !out->x || !out->y || !out->cx || !out->cy || !out->ok || !out->backward)
    return fail(TL_ERR_INVALID, "NULL output pointer%s");
  DeviceInfo info;
  rc = device_info(info);
  if (rc) return rc;
  const bool stacks = out->z_relu || out->theta || out->theta_prime;
  if (stacks && !(out->z_relu && out->theta && out->theta_prime))
    return fail(TL_ERR_INVALID, "the three aggregate=True stacks come together: z_relu, theta, theta_prime%s");
  if (stacks && is_general(*pb))
    return fail(TL_ERR_INVALID, "the aggregate=True stacks exist for spherical lenses only (not with k / a / sd)%s");
  // short pupil axis (the batched-lens workload: 64 rays per (lens, field, wavelength)): the
  // (pupil, wavelength)-flattened map keeps a CTA's threads busy where a CTA per row would idle
  const size_t smem_pw = (size_t)pb->W * ((5 * (size_t)pb->S + 3) & ~(size_t)3) * sizeof(float);
  // long rows of a spherical lens: the warp-owned forward kernel of spot_rev.cuh (any pupil map, any S <= 256)
  const bool rev_ok = !is_general(*pb) && !stacks && use_rev_eval_kernel(*pb);
  const bool many_short = pb->P < 2 * kFwdThreads && !getenv("TL_NO_ROWS");
  const bool short_rows = !is_general(*pb) && smem_pw <= 48 * 1024 && (many_short || ((pb->aim || pb->vig) && !rev_ok));
  if ((pb->aim || pb->vig) && !short_rows && !rev_ok)
    return fail(TL_ERR_INVALID, "an aimed forward trace needs W * S surface tables within 48 KB of shared memory%s");
  if (stacks || short_rows) {
    const int64_t row_len = (int64_t)pb->P * pb->W;
    if (row_len > 0x7fffffff) return fail(TL_ERR_INVALID, "P * W exceeds 2^31 - 1%s");
    const size_t smem = smem_pw;
    if (smem > 48 * 1024)
      return fail(TL_ERR_INVALID, "aggregate=True: W * S surface tables exceed 48 KB of shared memory%s");
    const FwdPlan pp = make_fwd_plan(info.sms, pb->B * pb->F, (int)row_len, kFwdThreads, 4);
    if (stacks)
      k_trace_fwd_pw<true><<<pp.n_blocks, kFwdThreads, smem, (cudaStream_t)stream_>>>(*pb, *out, pp.nchunks,
                                                                                     pp.chunk_len);
    else
      k_trace_fwd_pw<false><<<pp.n_blocks, kFwdThreads, smem, (cudaStream_t)stream_>>>(*pb, *out, pp.nchunks,
                                                                                      pp.chunk_len);
    g_launches++;
    TL_CHECK_CUDA(cudaGetLastError());
    return TL_OK;
  }
  const FwdPlan pl = make_fwd_plan(info.sms, pb->B * pb->F * pb->W, pb->P, 2 * kFwdThreads, 4);
  if (is_general(*pb)) {
    const size_t smem = gen_table_floats(pb->S) * sizeof(float);
    k_trace_fwd_gen<<<pl.n_blocks, kFwdThreads, smem, (cudaStream_t)stream_>>>(*pb, *out, pl.nchunks,
                                                                               pl.chunk_len);
    g_launches++;
    TL_CHECK_CUDA(cudaGetLastError());
    return TL_OK;
  }
  if (rev_ok) return launch_trace_rev(*pb, *out, (cudaStream_t)stream_);
  const size_t smem = 5 * (size_t)pb->S * sizeof(float);
  k_trace_fwd<<<pl.n_blocks, kFwdThreads, smem, (cudaStream_t)stream_>>>(*pb, *out, pl.nchunks,
                                                                         pl.chunk_len);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

size_t tl_trace_bwd_workspace(const TlProblem *pb) {
  if (validate(pb, TL_MAX_SURFACES_BWD)) return 0;
  if (is_general(*pb)) {
    TlProblem full = *pb;
    full.p_begin = 0;
    full.p_end = pb->P;
    GenPlan gp;
    if (plan_gen(full, 1, gp, true)) return 0;
    return gp.partial_bytes + align8((size_t)pb->B * pb->F * pb->W * gp.n_acc * sizeof(double));
  }
  TlProblem full = *pb;
  full.p_begin = 0;
  full.p_end = pb->P;
  AdjPlan pl, pl_pen;
  if (plan_adj(full, MODE_BWD, pl) || plan_adj(full, MODE_BWD, pl_pen, PEN_SEEDED)) return 0;
  const size_t partial = pl.partial_bytes > pl_pen.partial_bytes ? pl.partial_bytes : pl_pen.partial_bytes;
  return partial + align8((size_t)pb->B * pb->F * pb->W * pl.n_acc * sizeof(double));
}

int tl_trace_bwd(const TlProblem *pb_, const TlSeeds *seeds, const TlGrads *grads, void *workspace,
                 size_t workspace_bytes, void *stream_) {
  int rc = validate(pb_, TL_MAX_SURFACES_BWD);
  if (rc) return rc;
  if (!seeds || !grads || !grads->gc || !grads->gt || !grads->gmu || !grads->gz_sum)
    return fail(TL_ERR_INVALID, "NULL seeds/grads%s");
  if ((pb_->aim || pb_->vig) && (grads->gx || grads->gy))
    return fail(TL_ERR_INVALID, "per-ray gradients of x / y are not available through a ray-aiming map%s");
  if (is_general(*pb_)) {
    if (seeds->gz_relu || seeds->gtheta || seeds->gtheta_prime)
      return fail(TL_ERR_INVALID, "the aggregate=True stacks exist for spherical lenses only (no seeds on them with k / a / sd)%s");
    if (!grads->gk || !grads->ga) return fail(TL_ERR_INVALID, "general-surface lens: gk and ga are required%s");
    TlProblem pb = *pb_;
    pb.p_begin = 0;
    pb.p_end = pb.P;
    GenPlan gp;
    rc = plan_gen(pb, 1, gp, true);
    if (rc) return rc;
    const int rows = pb.B * pb.F * pb.W;
    const size_t rows_bytes = align8((size_t)rows * gp.n_acc * sizeof(double));
    if (!workspace || workspace_bytes < gp.partial_bytes + rows_bytes)
      return fail(TL_ERR_WORKSPACE, "workspace too small for tl_trace_bwd%s");
    cudaStream_t stream = (cudaStream_t)stream_;
    AdjArgs args;
    memset(&args, 0, sizeof(args));
    args.seeds = *seeds;
    args.grads = *grads;
    args.partial = (double *)workspace;
    args.groups_per_row = gp.groups_per_row;
    args.max_seg = gp.max_seg;
    args.n_acc = gp.n_acc;
    void *params[] = {(void *)&pb, (void *)&args};
    TL_CHECK_CUDA(cudaLaunchKernel((const void *)gp.kernel, dim3(gp.n_blocks), dim3(kTraceThreads), params,
                                   gp.smem, stream));
    g_launches++;
    double *rowbuf = (double *)((char *)workspace + gp.partial_bytes);
    const int64_t n = (int64_t)rows * gp.n_acc;
    k_reduce_rows<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(args.partial, rowbuf, rows,
                                                                   gp.groups_per_row, gp.n_blocks,
                                                                   gp.max_seg, gp.n_acc);
    g_launches++;
    const int outs = pb.B * (pb.S * (kGenPar - 1) + pb.W * pb.S + 1);
    k_bwd_finalize_gen<<<(outs + 127) / 128, 128, 0, stream>>>(rowbuf, *grads, grads->gk, grads->ga, pb.B,
                                                               pb.F, pb.W, pb.S);
    g_launches++;
    TL_CHECK_CUDA(cudaGetLastError());
    return TL_OK;
  }
  TlProblem pb = *pb_;
  pb.p_begin = 0;
  pb.p_end = pb.P;
  if (seeds->gopl)
    return fail(TL_ERR_INVALID, "the optical path length is an output of general-surface lenses only (no seed on it without k / a / sd)%s");
  const int pen = (seeds->gz_relu || seeds->gtheta || seeds->gtheta_prime) ? PEN_SEEDED : PEN_NONE;
  AdjPlan pl;
  rc = plan_adj(pb, MODE_BWD, pl, pen);
  if (rc) return rc;
  const int rows = pb.B * pb.F * pb.W;
  const size_t rows_bytes = align8((size_t)rows * pl.n_acc * sizeof(double));
  if (!workspace || workspace_bytes < pl.partial_bytes + rows_bytes)
    return fail(TL_ERR_WORKSPACE, "workspace too small for tl_trace_bwd%s");
  cudaStream_t stream = (cudaStream_t)stream_;
  AdjArgs args;
  memset(&args, 0, sizeof(args));
  args.seeds = *seeds;
  args.grads = *grads;
  args.partial = (double *)workspace;
  double *rowbuf = (double *)((char *)workspace + pl.partial_bytes);
  if (pb.S <= TL_MAX_SURFACES_SPOT && use_rows_kernel(pb, 1)) {     // many short rows: a warp per row
    rc = launch_bwd_rows(pb, args, pen, rowbuf, stream);
    if (rc) return rc;
  } else {
    rc = launch_adj(pb, args, pl, stream);
    if (rc) return rc;
    rc = reduce_rows(pl, args.partial, rowbuf, rows, stream);
    if (rc) return rc;
  }
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

int32_t tl_spot_moment_count_general(int32_t S, int32_t want_grad) { return n_acc_gen(S, want_grad); }

size_t tl_spot_workspace(const TlProblem *pb, int32_t want_grad) {
  if (validate(pb, want_grad ? TL_MAX_SURFACES_SPOT : TL_MAX_SURFACES_FWD)) return 0;
  if (pb->p_begin < 0 || pb->p_end > pb->P || pb->p_end <= pb->p_begin) return 0;
  if (is_general(*pb)) {
    GenPlan gp;
    if (plan_gen(*pb, want_grad, gp)) return 0;
    return gp.partial_bytes;
  }
  size_t bytes = 0;
  if (!(want_grad && pb->S > kLegacySpotSurfaces)) {
    AdjPlan pl;
    if (plan_adj(*pb, want_grad ? MODE_SPOT_GRAD : MODE_SPOT_EVAL, pl)) return 0;
    bytes = pl.partial_bytes;
  } else if (!use_rev_kernel(*pb)) {
    return 0;
  }
  if (want_grad ? use_rev_kernel(*pb) : use_rev_eval_kernel(*pb)) {
    RevPlan rp;
    if (plan_rev(*pb, rp, want_grad)) return 0;
    if (rp.partial_bytes > bytes) bytes = rp.partial_bytes;
  }
  return bytes;
}

static int spot_accumulate(const TlProblem *pb, int32_t want_grad, double *moments, float *ref_y, bool ref_ready,
                           void *workspace, size_t workspace_bytes, void *stream_);

int tl_spot_accumulate(const TlProblem *pb, int32_t want_grad, double *moments, float *ref_y,
                       void *workspace, size_t workspace_bytes, void *stream_) {
  return spot_accumulate(pb, want_grad, moments, ref_y, false, workspace, workspace_bytes, stream_);
}

int tl_spot_accumulate_ref(const TlProblem *pb, int32_t want_grad, double *moments, const float *ref_y,
                           void *workspace, size_t workspace_bytes, void *stream_) {
  return spot_accumulate(pb, want_grad, moments, const_cast<float *>(ref_y), true, workspace, workspace_bytes,
                         stream_);
}

static int spot_accumulate(const TlProblem *pb, int32_t want_grad, double *moments, float *ref_y, bool ref_ready,
                           void *workspace, size_t workspace_bytes, void *stream_) {
  int rc = validate(pb, want_grad ? TL_MAX_SURFACES_SPOT : TL_MAX_SURFACES_FWD);
  if (rc) return rc;
  if (pb->p_begin < 0 || pb->p_end > pb->P || pb->p_end <= pb->p_begin)
    return fail(TL_ERR_INVALID, "empty or out-of-range pupil slice%s");
  if (!moments || !ref_y) return fail(TL_ERR_INVALID, "NULL moments/ref_y%s");
  if (is_general(*pb)) {
    GenPlan gp;
    rc = plan_gen(*pb, want_grad, gp);
    if (rc) return rc;
    if (!workspace || workspace_bytes < gp.partial_bytes)
      return fail(TL_ERR_WORKSPACE, "workspace too small for tl_spot_accumulate%s");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n_bf = pb->B * pb->F;
    if (!ref_ready) {
      k_chief_rays_gen<<<(n_bf + 127) / 128, 128, 0, stream>>>(*pb, ref_y);
      g_launches++;
    }
    AdjArgs args;
    memset(&args, 0, sizeof(args));
    args.partial = (double *)workspace;
    args.ref_y = ref_y;
    args.groups_per_row = gp.groups_per_row;
    args.max_seg = gp.max_seg;
    args.n_acc = gp.n_acc;
    TlProblem pb_copy = *pb;
    void *params[] = {(void *)&pb_copy, (void *)&args};
    TL_CHECK_CUDA(cudaLaunchKernel((const void *)gp.kernel, dim3(gp.n_blocks), dim3(kTraceThreads),
                                   params, gp.smem, stream));
    g_launches++;
    const int rows = pb->B * pb->F * pb->W;
    const int64_t n = (int64_t)rows * gp.n_acc;
    k_reduce_rows<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(args.partial, moments, rows,
                                                                   gp.groups_per_row, gp.n_blocks,
                                                                   gp.max_seg, gp.n_acc);
    g_launches++;
    TL_CHECK_CUDA(cudaGetLastError());
    return TL_OK;
  }
  const int mode = want_grad ? MODE_SPOT_GRAD : MODE_SPOT_EVAL;
  const bool legacy_fits = !(want_grad && pb->S > kLegacySpotSurfaces);
  if (!legacy_fits && !use_rev_kernel(*pb))
    return fail(TL_ERR_INVALID, "the fused pass with gradients beyond 16 surfaces needs k_spot_rev (TL_NO_REV is set)%s");
  AdjPlan pl;
  if (legacy_fits) {
    rc = plan_adj(*pb, mode, pl);
    if (rc) return rc;
    if (!workspace || workspace_bytes < pl.partial_bytes)
      return fail(TL_ERR_WORKSPACE, "workspace too small for tl_spot_accumulate%s");
  }
  if (!workspace) return fail(TL_ERR_WORKSPACE, "workspace too small for tl_spot_accumulate%s");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int n_bf = pb->B * pb->F;
  if (!ref_ready) {
    k_chief_rays<<<(n_bf + 127) / 128, 128, 0, stream>>>(*pb, ref_y);
    g_launches++;
  }
  if (use_rows_kernel(*pb, want_grad))      // many short rows: a warp per row, sums straight into `moments`
    return launch_spot_rows(*pb, want_grad, ref_y, moments, stream);
  if (want_grad ? use_rev_kernel(*pb) : use_rev_eval_kernel(*pb)) {   // the reversible fused pass / its forward-only variant (spot_rev.cuh)
    RevPlan rp;
    rc = plan_rev(*pb, rp, want_grad);
    if (rc) return rc;
    if (workspace_bytes < rp.partial_bytes)
      return fail(TL_ERR_WORKSPACE, "workspace too small for tl_spot_accumulate%s");
    return launch_spot_rev(*pb, rp, ref_y, (double *)workspace, moments, stream);
  }
  AdjArgs args;
  memset(&args, 0, sizeof(args));
  args.partial = (double *)workspace;
  args.ref_y = ref_y;
  rc = launch_adj(*pb, args, pl, stream);
  if (rc) return rc;
  return reduce_rows(pl, args.partial, moments, pb->B * pb->F * pb->W, stream);
}

const char *tl_spot_kernel_name(const TlProblem *pb, int32_t want_grad) {
  static thread_local char text[96];
  if (validate(pb, want_grad ? TL_MAX_SURFACES_SPOT : TL_MAX_SURFACES_FWD)) return "invalid";
  if (is_general(*pb)) return want_grad ? "k_trace_gen<SPOT_GRAD,f2>" : "k_trace_gen<SPOT_EVAL,f4>";
  if (use_rows_kernel(*pb, want_grad)) return want_grad ? "k_spot_rows<GRAD,f2>" : "k_spot_rows<EVAL,f2>";
  if (want_grad ? use_rev_kernel(*pb) : use_rev_eval_kernel(*pb)) {
    RevPlan rp;
    if (plan_rev(*pb, rp, want_grad)) return "invalid";
    snprintf(text, sizeof(text), "k_spot_rev<%s,f4>", rp.name);
    return text;
  }
  return want_grad ? "k_trace_adj<SPOT_GRAD>" : "k_trace_adj<SPOT_EVAL,f4>";
}

// Diagnostics for bench.py's roofline line: the dominant kernel of the fused pass ALONE (the chief-ray
// and row-reduction launches that tl_spot_accumulate adds around it are left out), so that CUDA
// events around this call time exactly one kernel.  `ref_y` must hold the reference heights.
int tl_spot_kernel_only(const TlProblem *pb, const float *ref_y, void *workspace, size_t workspace_bytes,
                        void *stream_) {
  int rc = validate(pb, TL_MAX_SURFACES_SPOT);
  if (rc) return rc;
  if (!ref_y || !workspace) return fail(TL_ERR_INVALID, "tl_spot_kernel_only: NULL argument%s");
  if (is_general(*pb) || use_rows_kernel(*pb, 1) || !use_rev_kernel(*pb))
    return fail(TL_ERR_INVALID, "tl_spot_kernel_only: this problem does not take k_spot_rev%s");
  RevPlan pl;
  rc = plan_rev(*pb, pl);
  if (rc) return rc;
  if (workspace_bytes < pl.partial_bytes) return fail(TL_ERR_WORKSPACE, "workspace too small%s");
  RevArgs args{};
  args.partial = (double *)workspace;
  args.ref_y = ref_y;
  args.groups_per_row = pl.groups_per_row;
  args.max_owners = pl.max_owners;
  args.n_acc = pl.n_acc;
  TlProblem pb_copy = *pb;
  void *params[] = {(void *)&pb_copy, (void *)&args};
  TL_CHECK_CUDA(cudaLaunchKernel((const void *)pl.kernel, dim3(pl.n_blocks), dim3(pl.n_warps_cta * 32), params,
                                 pl.smem, (cudaStream_t)stream_));
  g_launches++;
  return TL_OK;
}

static int check_paraxial(const TlParaxial *p, const char *who) {
  if (!p || !p->c || !p->t || !p->n || !p->live || !p->glass) return fail(TL_ERR_INVALID, "tl_paraxial: NULL argument%s");
  if (p->B < 1 || p->L < 1 || p->L > TL_PARAXIAL_MAX_SLOTS)
    return fail(TL_ERR_INVALID, "tl_paraxial: B >= 1 and 1 <= L <= 64%s");
  if (p->mode != TL_PARAXIAL_FIRST_ORDER && p->mode != TL_PARAXIAL_LAST_CURVATURE)
    return fail(TL_ERR_INVALID, "tl_paraxial: unknown mode%s");
  (void)who;
  return TL_OK;
}

int tl_paraxial_fwd(const TlParaxial *lens, float *out, void *stream_) {
  int rc = check_paraxial(lens, "tl_paraxial_fwd");
  if (rc) return rc;
  if (!out) return fail(TL_ERR_INVALID, "tl_paraxial_fwd: NULL output%s");
  k_paraxial_fwd<<<(lens->B + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(*lens, out);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

int tl_paraxial_bwd(const TlParaxial *lens, const float *gout, float *gc, float *gt, float *gn, void *stream_) {
  int rc = check_paraxial(lens, "tl_paraxial_bwd");
  if (rc) return rc;
  if (!gout || !gc || !gt || !gn) return fail(TL_ERR_INVALID, "tl_paraxial_bwd: NULL argument%s");
  k_paraxial_bwd<<<(lens->B + 63) / 64, 64, 0, (cudaStream_t)stream_>>>(*lens, gout, gc, gt, gn);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

size_t tl_psf_workspace(const TlPsf *psf) {
  PsfPlan pl;
  if (!psf || plan_psf(*psf, pl)) return 0;
  return pl.partial_bytes;
}

int tl_psf_bin(const TlPsf *psf, double *sums, double *inside, void *workspace, size_t workspace_bytes,
               void *stream_) {
  if (!psf || !psf->x || !psf->y || !psf->y_target || !psf->x_incr || !psf->y_incr || !psf->x_size ||
      !psf->y_size || !sums || !inside)
    return fail(TL_ERR_INVALID, "tl_psf_bin: NULL argument%s");
  PsfPlan pl;
  int rc = plan_psf(*psf, pl);
  if (rc) return rc;
  if (!workspace || workspace_bytes < pl.partial_bytes)
    return fail(TL_ERR_WORKSPACE, "workspace too small for tl_psf_bin%s");
  if (pl.smem > 48 * 1024)
    TL_CHECK_CUDA(cudaFuncSetAttribute((const void *)k_psf_bin, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)pl.smem));
  cudaStream_t stream = (cudaStream_t)stream_;
  PsfArgs args;
  args.p = *psf;
  args.partial = (double *)workspace;
  args.n_chunks = pl.n_chunks;
  args.chunk_len = pl.chunk_len;
  args.n_xh = pl.n_xh;
  const int n_gc = psf->G * psf->C, n = pl.n_xh * psf->n_y_bins + 1;
  k_psf_bin<<<n_gc * pl.n_chunks, kPsfThreads, pl.smem, stream>>>(args);
  g_launches++;
  k_psf_reduce<<<(unsigned)(((int64_t)n_gc * n + 255) / 256), 256, 0, stream>>>(args.partial, sums, inside, n_gc,
                                                                                  pl.n_chunks, n);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

int32_t tl_penalty_moment_count(int32_t S) { return 3 * S + 2; }

size_t tl_penalty_workspace(const TlProblem *pb) {
  if (validate(pb, TL_MAX_SURFACES_BWD) || is_general(*pb)) return 0;
  if (pb->p_begin < 0 || pb->p_end > pb->P || pb->p_end <= pb->p_begin) return 0;
  AdjPlan pl;
  if (plan_adj(*pb, MODE_BWD, pl, PEN_SUM)) return 0;
  return pl.partial_bytes;
}

int tl_penalty_accumulate(const TlProblem *pb, double *moments, void *workspace, size_t workspace_bytes,
                          void *stream_) {
  int rc = validate(pb, TL_MAX_SURFACES_BWD);
  if (rc) return rc;
  if (is_general(*pb))
    return fail(TL_ERR_INVALID, "the penalty terms exist for spherical lenses only (not with k / a / sd)%s");
  if (pb->p_begin < 0 || pb->p_end > pb->P || pb->p_end <= pb->p_begin)
    return fail(TL_ERR_INVALID, "empty or out-of-range pupil slice%s");
  if (!moments) return fail(TL_ERR_INVALID, "NULL moments%s");
  if (pb->S <= TL_MAX_SURFACES_SPOT && use_rows_kernel(*pb, 1)) {   // many short rows: a warp per row
    AdjArgs none;
    memset(&none, 0, sizeof(none));
    return launch_bwd_rows(*pb, none, PEN_SUM, moments, (cudaStream_t)stream_);
  }
  AdjPlan pl;
  rc = plan_adj(*pb, MODE_BWD, pl, PEN_SUM);
  if (rc) return rc;
  if (!workspace || workspace_bytes < pl.partial_bytes)
    return fail(TL_ERR_WORKSPACE, "workspace too small for tl_penalty_accumulate%s");
  cudaStream_t stream = (cudaStream_t)stream_;
  AdjArgs args;
  memset(&args, 0, sizeof(args));
  args.partial = (double *)workspace;
  rc = launch_adj(*pb, args, pl, stream);
  if (rc) return rc;
  return reduce_rows(pl, args.partial, moments, pb->B * pb->F * pb->W, stream);
}

int tl_penalty_finalize(const double *moments, int32_t B, int32_t F, int32_t W, int32_t S, double scale,
                        const TlPenaltyOut *out, void *stream_) {
  if (!moments || !out || !out->penalty || !out->gc || !out->gt || !out->gmu || !out->gz || B < 1 ||
      F < 1 || W < 1 || S < 1 || S > TL_MAX_SURFACES_BWD)
    return fail(TL_ERR_INVALID, "bad argument to tl_penalty_finalize%s");
  const int outs = B * (2 * S + W * S + 2);
  k_penalty_finalize<<<(outs + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(moments, *out, B, F, W, S, scale);
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
  if ((size_t)F * 3 * sizeof(double) > 40 * 1024)
    return fail(TL_ERR_INVALID, "too many fields%s");
  if (out->gk || out->ga) {                      // general-surface moment layout
    if (want_grad && (!out->gk || !out->ga)) return fail(TL_ERR_INVALID, "gk and ga go together%s");
    k_spot_finalize_gen<<<B, 256, (size_t)F * 3 * sizeof(double), (cudaStream_t)stream_>>>(
        moments, ref_y, B, F, W, S, (double)P_total * (double)W, want_grad, *out);
    g_launches++;
    TL_CHECK_CUDA(cudaGetLastError());
    return TL_OK;
  }
  const int n_acc = n_acc_of(want_grad ? MODE_SPOT_GRAD : MODE_SPOT_EVAL, S);
  const size_t base = (size_t)F * 3 * sizeof(double);
  const size_t staged = base + (size_t)F * W * n_acc * sizeof(double);
  const double n_rays = (double)P_total * (double)W;
  if (staged <= 160 * 1024) {
    if (staged > 48 * 1024)
      TL_CHECK_CUDA(cudaFuncSetAttribute((const void *)k_spot_finalize<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged));
    k_spot_finalize<true><<<B, 256, staged, (cudaStream_t)stream_>>>(moments, ref_y, B, F, W, S, n_rays,
                                                                     want_grad, *out);
  } else {
    k_spot_finalize<false><<<B, 256, base, (cudaStream_t)stream_>>>(moments, ref_y, B, F, W, S, n_rays,
                                                                    want_grad, *out);
  }
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

static int validate_lens(const TlLens *ln) {
  if (!ln || !ln->c || !ln->t || !ln->nd || !ln->v || !ln->mask || !ln->mask_g || !ln->stop_idx ||
      !ln->hfov || !ln->epd || !ln->rel_fields || !ln->wavelengths)
    return fail(TL_ERR_INVALID, "NULL lens field%s");
  if (ln->B < 1 || ln->L < 1 || ln->F < 1 || ln->W < 1 || ln->L > kStageMaxSurfaces)
    return fail(TL_ERR_INVALID, "bad lens sizes (B, L, F, W >= 1, L <= 64)%s");
  return TL_OK;
}

int tl_stage_fwd(const TlLens *lens, float *mu, float *z, float *cy, float *half_epd, void *stream_) {
  int rc = validate_lens(lens);
  if (rc) return rc;
  if (!mu || !z || !cy || !half_epd) return fail(TL_ERR_INVALID, "NULL output of tl_stage_fwd%s");
  k_stage_fwd<<<lens->B, 128, 0, (cudaStream_t)stream_>>>(*lens, mu, z, cy, half_epd);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

int tl_stage_ref(const TlLens *lens, const TlProblem *pb, float *mu, float *z, float *cy, float *half_epd,
                 float *aim, const float *vig, int32_t aim_mode, int32_t allow_backward_rays, float *ref_y,
                 void *stream_) {
  int rc = validate_lens(lens);
  if (rc) return rc;
  if (!mu || !z || !cy || !half_epd) return fail(TL_ERR_INVALID, "NULL output of tl_stage_ref%s");
  TlProblem none;
  memset(&none, 0, sizeof(none));
  if (ref_y) {
    rc = validate(pb, TL_MAX_SURFACES_FWD);
    if (rc) return rc;
    if (is_general(*pb)) return fail(TL_ERR_INVALID, "tl_stage_ref: spherical lenses only%s");
    if (pb->B != lens->B || pb->F != lens->F || pb->W != lens->W || pb->S != lens->L)
      return fail(TL_ERR_INVALID, "tl_stage_ref: problem and lens sizes differ%s");
  }
  if (aim_mode != TL_AIM_REAL && aim_mode != TL_AIM_PARAXIAL) return fail(TL_ERR_INVALID, "unknown aim_mode%s");
  k_stage_ref<<<lens->B, 128, 0, (cudaStream_t)stream_>>>(*lens, ref_y ? *pb : none, mu, z, cy, half_epd, aim, vig,
                                                          aim_mode, allow_backward_rays, ref_y);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

int tl_lens_spot_finalize(const double *moments, const float *ref_y, const TlLens *lens, int64_t P_total,
                          const TlSpotOut *out, float *gnd, float *gv, void *stream_) {
  int rc = validate_lens(lens);
  if (rc) return rc;
  if (!moments || !ref_y || !out || !out->rms || !out->rms_field || !out->gc || !out->gt || !out->gmu ||
      !out->gz || !gnd || !gv || P_total < 1)
    return fail(TL_ERR_INVALID, "bad argument to tl_lens_spot_finalize%s");
  const int B = lens->B, F = lens->F, W = lens->W, S = lens->L;
  if (S > TL_MAX_SURFACES_SPOT) return fail(TL_ERR_INVALID, "too many surfaces for tl_lens_spot_finalize%s");
  if ((size_t)F * 3 * sizeof(double) > 40 * 1024) return fail(TL_ERR_INVALID, "too many fields%s");
  const int n_acc = n_acc_of(MODE_SPOT_GRAD, S);
  const size_t base = (size_t)F * 3 * sizeof(double);
  const size_t staged = base + (size_t)F * W * n_acc * sizeof(double);
  const double n_rays = (double)P_total * (double)W;
  if (staged <= 160 * 1024) {
    if (staged > 48 * 1024)
      TL_CHECK_CUDA(cudaFuncSetAttribute((const void *)k_lens_finalize<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged));
    k_lens_finalize<true><<<B, 256, staged, (cudaStream_t)stream_>>>(moments, ref_y, B, F, W, S, n_rays, *out, *lens,
                                                                     gnd, gv);
  } else {
    k_lens_finalize<false><<<B, 256, base, (cudaStream_t)stream_>>>(moments, ref_y, B, F, W, S, n_rays, *out, *lens,
                                                                    gnd, gv);
  }
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

int tl_aim(const TlLens *lens, const float *mu, const float *z, const float *cy, const float *half_epd,
           const float *vig, int32_t aim_mode, int32_t allow_backward_rays, float *aim, void *stream_) {
  int rc = validate_lens(lens);
  if (rc) return rc;
  if (!mu || !z || !cy || !half_epd || !aim) return fail(TL_ERR_INVALID, "NULL argument of tl_aim%s");
  const int n = lens->B * lens->F * lens->W;
  if (aim_mode != TL_AIM_REAL && aim_mode != TL_AIM_PARAXIAL) return fail(TL_ERR_INVALID, "unknown aim_mode%s");
  k_aim<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(*lens, mu, z, cy, half_epd, vig, aim_mode,
                                                           allow_backward_rays, aim);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

int tl_stage_bwd(const TlLens *lens, const float *gmu, const float *gz, float *gc, float *gt,
                 float *gnd, float *gv, void *stream_) {
  int rc = validate_lens(lens);
  if (rc) return rc;
  if (!gmu || !gz || !gc || !gt || !gnd || !gv) return fail(TL_ERR_INVALID, "NULL argument of tl_stage_bwd%s");
  k_stage_bwd<<<lens->B, 64, 0, (cudaStream_t)stream_>>>(*lens, gmu, gz, gc, gt, gnd, gv);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

}  // extern "C"

#include "peer_exchange.cuh"
