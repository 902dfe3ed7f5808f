// Kernels for lenses with EXTENSION surfaces (conic / even asphere / clear semi-diameter,
// optical path length): included by trace_kernels.cu inside its anonymous namespace.
//
// Same structure as the spherical kernels (persistent CTAs over contiguous (row, group)
// slices, packed two-ray lanes, guarded fast path + exact re-trace, parked hit point +
// incoming direction, geometric adjoint sweep), with two differences forced by the 11
// gradients per surface (c, k, a4..a16, t, mu), twice (plain and (y - y0)-weighted):
//   * the accumulators do not fit the register file, so every adjoint step reduces its 22
//     values across the warp with a halving butterfly (24 shuffles; every lane ends up with the
//     warp total of one of them) and adds them to a per-warp accumulator row in shared memory;
//   * the surface loops are therefore rolled (no static accumulator indices needed).
#pragma once

constexpr int kGenPar = kAsphParams + 2;        // c, k, a4..a16, t, mu  = 11 per surface
constexpr int kGenSlots = 2 * kGenPar;           // weighted + plain      = 22 per surface
constexpr int kGenRow = 32;                      // padded row of the per-warp accumulators

// Surface table of one (lens, wavelength) in shared memory, general surfaces: ONE 64-byte record per
// surface, so that a step of either loop reads it with four 128-bit broadcast loads from one address
// (the first version kept eight separate arrays: 15 scalar loads and their address arithmetic per
// surface and pass, a tenth of the fused pass's instructions).
struct alignas(16) GenSurf {
  float c, k, t, mu;
  float sd2, index, index_next;     // sd2 = (clear semi-diameter)^2; refractive index in front of / behind the surface
  int live_bits;                    // bit 0: structure mask of this surface, bit 1: ... of the surface in front
  float a[kAsphCoefs];
  float pad;
};
static_assert(sizeof(GenSurf) == 64, "GenSurf is read as four float4");

struct GenTable {
  const GenSurf *s;
  float length;                     // sum |t|
  float index_image, t_last;        // index behind the last surface; its thickness
  int live_last;
};

__host__ __device__ __forceinline__ size_t gen_table_floats(int S) { return 16 * (size_t)S; }

__device__ __forceinline__ GenTable load_gen_table(float *base, const TlProblem &pb, int b, int w) {
  GenTable tab;
  const int S = pb.S;
  GenSurf *rec = reinterpret_cast<GenSurf *>(base);
  tab.s = rec;
  for (int k = threadIdx.x; k < S; k += blockDim.x) {
    const int64_t i = (int64_t)b * S + k;
    GenSurf r;
    r.c = pb.c[i];
    r.k = pb.k ? pb.k[i] : 0.f;
    r.t = pb.t[i];
    r.mu = pb.mu[((int64_t)b * pb.W + w) * S + k];
    const float sd = pb.sd ? pb.sd[i] : INFINITY;
    r.sd2 = __fmul_rn(sd, sd);
    r.index = r.index_next = 1.0f;             // (filled below)
    r.live_bits = (pb.live[i] != 0 ? 1 : 0) | (k > 0 && pb.live[i - 1] != 0 ? 2 : 0);
    for (int j = 0; j < kAsphCoefs; ++j) r.a[j] = pb.a ? pb.a[i * kAsphCoefs + j] : 0.f;
    r.pad = 0.f;
    rec[k] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float n = 1.0f;
    for (int k = 0; k < S; ++k) {
      rec[k].index = n;
      n = __fdiv_rn(n, rec[k].mu);             // like the oracle: index = index / mu
      rec[k].index_next = n;
    }
  }
  __syncthreads();
  float len = 0.f;
  for (int k = 0; k < S; ++k) len += fabsf(rec[k].t);
  tab.length = len;
  tab.index_image = rec[S - 1].index_next;
  tab.t_last = rec[S - 1].t;
  tab.live_last = rec[S - 1].live_bits & 1;
  return tab;
}

// one record, as the kernels use it
struct GenStep {
  AsphSurface s;
  float index, index_next;
  bool live_prev;
};

__device__ __forceinline__ GenStep gen_step(const GenTable &tab, int k) {
  const float4 *p = reinterpret_cast<const float4 *>(tab.s + k);
  const float4 q0 = p[0], q1 = p[1], q2 = p[2], q3 = p[3];
  GenStep g;
  g.s.c = q0.x;
  g.s.k = q0.y;
  g.s.t = q0.z;
  g.s.mu = q0.w;
  g.s.sd2 = q1.x;
  g.index = q1.y;
  g.index_next = q1.z;
  g.live_prev = (__float_as_int(q1.w) & 2) != 0;
  g.s.a[0] = q2.x;
  g.s.a[1] = q2.y;
  g.s.a[2] = q2.z;
  g.s.a[3] = q2.w;
  g.s.a[4] = q3.x;
  g.s.a[5] = q3.y;
  g.s.a[6] = q3.z;
  return g;
}

struct TracedGen {
  Ray<float> pre;
  float x, y, opl;
  bool ok, backward;
};

template <bool SAVE>
__device__ __noinline__ TracedGen trace_exact_gen(float x, float y, float z, float cx, float cy,
                                                  const GenTable &tab, int S, bool allow_backward,
                                                  float *state, int stride) {
  Ray<float> r{x, y, z, cx, cy, exact_cz0(cx, cy)};
  bool ok = true, backward = false;
  float index = 1.0f, opl = 0.0f;
  for (int k = 0; k < S; ++k) {
    const float in_cx = r.cx, in_cy = r.cy;
    const GenStep st = gen_step(tab, k);
    exact_asph_surface(r, st.s, st.live_prev, allow_backward, ok, backward, index, opl);
    park<SAVE, float>(state, stride, k, r.x, r.y, in_cx, in_cy);
  }
  TracedGen out;
  out.pre = r;
  exact_asph_image(r, tab.live_last != 0, allow_backward, ok, backward, index, opl);
  out.x = r.x;
  out.y = r.y;
  out.opl = opl;
  out.ok = ok;
  out.backward = backward;
  return out;
}

template <class V>
struct TracedGenN {
  Ray<V> pre;
  V x, y, opl;
  bool ok[LaneCount<V>::value], backward[LaneCount<V>::value];
};

template <bool SAVE, class V>
__device__ __forceinline__ TracedGenN<V> trace_guarded_gen(V x, V y, V z, V cx, V cy,
                                                           const GenTable &tab, int S,
                                                           bool allow_backward, int arith, V *state,
                                                           int stride) {
  constexpr int N = LaneCount<V>::value;
  TracedGenN<V> out;
  bool clear[N];
#pragma unroll
  for (int l = 0; l < N; ++l) clear[l] = false;
  if (arith == TL_ARITH_GUARDED) {
    Ray<V> r{x, y, z, cx, cy, fast_cz0(cx, cy)};
    V min_cos2(1.0f), min_travel(3.0e38f), min_clip(3.0e38f), opl(0.0f);
#pragma unroll 1
    for (int k = 0; k < S; ++k) {
      const V in_cx = r.cx, in_cy = r.cy;
      V travel;
      // clip margin relative to the size of rho: computed against sd2 inside
      const GenStep st = gen_step(tab, k);
      fast_asph_surface(r, st.s, min_cos2, travel, min_clip, V(st.index), opl);
      park<SAVE, V>(state, stride, k, r.x, r.y, in_cx, in_cy);
      if (st.live_prev) min_travel = fmin2(min_travel, travel);
    }
    out.pre = r;
    const V rcz = frcp(r.cz);
    const V dist = -r.z * rcz;
    opl = ffma(V(tab.index_image), dist, opl);
    const V travel = fast_image(r);
    if (tab.live_last) min_travel = fmin2(min_travel, travel);
    out.x = r.x;
    out.y = r.y;
    out.opl = opl;
    const V probe = ((r.x + r.y) + (r.cx + r.cy)) + opl;
#pragma unroll
    for (int l = 0; l < N; ++l) {
      const float band = kBandTravelRel * fmaxf(1.0f, tab.length + fabsf(lane_get(z, l)));
      clear[l] = (lane_get(min_cos2, l) > kGuard + kBandCos2) && (lane_get(min_travel, l) > band) &&
                 (lane_get(min_clip, l) > kBandTravelRel * fmaxf(1.0f, tab.length * tab.length)) &&
                 (fabsf(lane_get(probe, l)) < 3.0e38f);
    }
  }
#pragma unroll
  for (int l = 0; l < N; ++l) {
    out.ok[l] = true;
    out.backward[l] = false;
    if (!clear[l]) {
      const TracedGen one = trace_exact_gen<SAVE>(lane_get(x, l), lane_get(y, l), lane_get(z, l),
                                                  lane_get(cx, l), lane_get(cy, l), tab, S,
                                                  allow_backward, reinterpret_cast<float *>(state) + l,
                                                  N * stride);
      lane_set(out.pre.x, l, one.pre.x);
      lane_set(out.pre.y, l, one.pre.y);
      lane_set(out.pre.z, l, one.pre.z);
      lane_set(out.pre.cx, l, one.pre.cx);
      lane_set(out.pre.cy, l, one.pre.cy);
      lane_set(out.pre.cz, l, one.pre.cz);
      lane_set(out.x, l, one.x);
      lane_set(out.y, l, one.y);
      lane_set(out.opl, l, one.opl);
      out.ok[l] = one.ok;
      out.backward[l] = one.backward;
    }
  }
  return out;
}

// Halving butterfly: v[0..31] per lane in, lane i holds sum over the warp of v[i] out (in v[0]).
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? v[i] : v[i + n / 2];
      const float keep = upper ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// The same for the general-surface pass: every lane brings the kGenSlots = 22 values of its rays (padded to 24), and
// after five exchange-and-add steps every lane holds the warp total of ONE of them.  Slot counts per lane
// 24 -> 12 -> 6 -> 3 (padded to 4) -> 2 -> 1: 12 + 6 + 3 + 2 + 1 = 24 shuffles (the first version padded to 32
// slots: 31 shuffles, 62 selects).  Which slot a lane ends with follows from the halves it kept:
// slot = 12 b4 + 6 b3 + 3 b2 + (2 b1 + b0) for lane bits b4..b0, lanes with b1 = b0 = 1 hold padding.
constexpr int kGenPad = 24;
static_assert(kGenSlots <= kGenPad, "the butterfly carries 24 slots");
template <int N>
__device__ __forceinline__ void halve_and_add(float (&v)[kGenPad], bool upper, int off) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float send = upper ? v[i] : v[i + N / 2];
    const float keep = upper ? v[i + N / 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
  }
}
__device__ __forceinline__ float warp_transpose_sum24(float (&v)[kGenPad], int lane) {
  halve_and_add<24>(v, (lane & 16) != 0, 16);
  halve_and_add<12>(v, (lane & 8) != 0, 8);
  halve_and_add<6>(v, (lane & 4) != 0, 4);
  v[3] = 0.f;
  halve_and_add<4>(v, (lane & 2) != 0, 2);
  halve_and_add<2>(v, (lane & 1) != 0, 1);
  return v[0];
}
// the lane that ends up with slot j (j < 22)
__host__ __device__ __forceinline__ int gen_lane_of_slot(int j) {
  return 16 * (j / 12) + 8 * ((j % 12) / 6) + 4 * ((j % 6) / 3) + j % 3;
}

// --------------------------------------------------------------------------
// forward trace, general surfaces (+ optical path length)
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(kFwdThreads)
k_trace_fwd_gen(TlProblem pb, TlTraceOut out, int nchunks, int chunk_len) {
  extern __shared__ __align__(16) float smem[];
  int blk = blockIdx.x;
  const int chunk = blk % nchunks; blk /= nchunks;
  const int w = blk % pb.W; blk /= pb.W;
  const int f = blk % pb.F;
  const int b = blk / pb.F;
  const GenTable tab = load_gen_table(smem, pb, b, w);
  const int p_lo = chunk * chunk_len;
  const int p_hi = min(pb.P, p_lo + chunk_len);
  const float xy_scale = pb.xy_scale ? pb.xy_scale[b] : 1.0f;
  for (int p0 = p_lo + threadIdx.x; p0 < p_hi; p0 += 2 * kFwdThreads) {
    const int p1 = p0 + kFwdThreads;
    const bool has1 = p1 < p_hi;
    const int q1 = has1 ? p1 : p0;
    f2 x(__fmul_rn(pb.x.ptr[offset_of(pb.x, b, f, p0, w)], xy_scale),
         __fmul_rn(pb.x.ptr[offset_of(pb.x, b, f, q1, w)], xy_scale));
    f2 y(__fmul_rn(pb.y.ptr[offset_of(pb.y, b, f, p0, w)], xy_scale),
         __fmul_rn(pb.y.ptr[offset_of(pb.y, b, f, q1, w)], xy_scale));
    const f2 z(pb.z.ptr[offset_of(pb.z, b, f, p0, w)], pb.z.ptr[offset_of(pb.z, b, f, q1, w)]);
    const f2 cx(pb.cx.ptr[offset_of(pb.cx, b, f, p0, w)], pb.cx.ptr[offset_of(pb.cx, b, f, q1, w)]);
    const f2 cy(pb.cy.ptr[offset_of(pb.cy, b, f, p0, w)], pb.cy.ptr[offset_of(pb.cy, b, f, q1, w)]);
    const TracedGenN<f2> tr = trace_guarded_gen<false, f2>(x, y, z, cx, cy, tab, pb.S,
                                                           pb.allow_backward_rays != 0, pb.arith,
                                                           nullptr, 0);
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
      if (out.opl) out.opl[o] = lane_get(tr.opl, l);
    }
  }
}

// Chief-ray reference heights for general lenses (fast policy, see k_chief_rays).
__global__ void k_chief_rays_gen(TlProblem pb, float *ref_y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pb.B * pb.F) return;
  const int b = i / pb.F, f = i % pb.F, S = pb.S;
  const float cx = pb.cx.ptr[offset_of(pb.cx, b, f, 0, 0)], cy = pb.cy.ptr[offset_of(pb.cy, b, f, 0, 0)];
  float x0, y0;      // the first ray of the bundle (see k_chief_rays)
  load_pupil_point<false>(pb, b, f, 0, 0, pb.xy_scale ? pb.xy_scale[b] : 1.0f, x0, y0);
  Ray<float> r{x0, y0, pb.z.ptr[offset_of(pb.z, b, f, 0, 0)], cx, cy, fast_cz0(cx, cy)};
  float min_cos2 = 1.0f, travel, min_clip = 3.0e38f, opl = 0.f;
  for (int k = 0; k < S; ++k) {
    const int64_t j = (int64_t)b * S + k;
    AsphSurface s;
    s.c = pb.c[j];
    s.t = pb.t[j];
    s.mu = pb.mu[((int64_t)b * pb.W) * S + k];
    s.k = pb.k ? pb.k[j] : 0.f;
    s.sd2 = INFINITY;
    for (int q = 0; q < kAsphCoefs; ++q) s.a[q] = pb.a ? pb.a[j * kAsphCoefs + q] : 0.f;
    fast_asph_surface(r, s, min_cos2, travel, min_clip, 1.0f, opl);
  }
  fast_image(r);
  ref_y[i] = (min_cos2 > kGuard && fabsf(r.y) < 3.0e38f) ? r.y : 0.f;
}

// --------------------------------------------------------------------------
// fused spot pass (MODE_SPOT_GRAD / MODE_SPOT_EVAL) and split backward (MODE_BWD), general
// surfaces.  Row layout: per surface 22 = [11 weighted | 11 plain] (c, k, a4..a16, t, mu), then
// {weighted z, plain z, S1, S2, n_ok}: n_acc = 22 S + 5  (EVAL: 3).  MODE_BWD uses the same
// layout with the caller's seeds as the only weights (the "plain" half stays zero).
// --------------------------------------------------------------------------
template <int MODE, class V>
__global__ void __launch_bounds__(kTraceThreads)
k_trace_gen(TlProblem pb, AdjArgs args) {
  extern __shared__ __align__(16) float smem[];
  constexpr int N = LaneCount<V>::value;
  constexpr bool kAdjoint = MODE != MODE_SPOT_EVAL;
  constexpr bool kSeeded = MODE == MODE_BWD;
  constexpr int kWarps = kTraceThreads / 32;
  const int S = pb.S;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int stride = kTraceThreads;
  const bool allow_backward = pb.allow_backward_rays != 0;
  const int n_acc = args.n_acc;
  float *after_table = smem + ((gen_table_floats(S) + 3) & ~(size_t)3);
  // [warp][surface][32] accumulator rows, then the parked states
  float *acc_rows = after_table;
  const size_t acc_floats = kAdjoint ? (size_t)kWarps * S * kGenRow : 0;
  V *state = reinterpret_cast<V *>(after_table + acc_floats) + tid;
  float *my_rows = acc_rows + (size_t)warp * S * kGenRow;

  const int64_t total = (int64_t)pb.B * pb.F * pb.W * args.groups_per_row;
  const int64_t g_begin = total * blockIdx.x / gridDim.x;
  const int64_t g_end = total * (blockIdx.x + 1) / gridDim.x;

  __shared__ float tail[kTraceThreads / 32][5];
  float acc_z = 0.f, wac_z = 0.f, m_s1 = 0.f, m_s2 = 0.f, m_n = 0.f;
  GenTable tab;
  float y0 = 0.f, xy_scale = 1.0f;
  int row = -1, seg = 0, b = 0, f = 0, w = 0;

  auto flush = [&]() {
    __syncthreads();
    double *dst = args.partial + ((int64_t)blockIdx.x * args.max_seg + seg) * n_acc;
    if (kAdjoint) {
      for (int i = tid; i < S * kGenSlots; i += kTraceThreads) {
        const int k = i / kGenSlots, j = i % kGenSlots;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < kWarps; ++q) s += (double)acc_rows[((size_t)q * S + k) * kGenRow + gen_lane_of_slot(j)];
        dst[i] = s;
      }
    }
    // the five per-thread scalars: warp shuffle -> smem -> fp64
    const float vals[5] = {wac_z, acc_z, m_s1, m_s2, m_n};
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float v = warp_sum(vals[j]);
      if (lane == 0) tail[warp][j] = v;
    }
    __syncthreads();
    const int base = kAdjoint ? S * kGenSlots : 0;
    const int first = kAdjoint ? 0 : 2;          // EVAL keeps only S1, S2, n_ok
    if (tid < 5 - first) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < kWarps; ++q) s += (double)tail[q][first + tid];
      dst[base + tid] = s;
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
      tab = load_gen_table(smem, pb, b, w);
      xy_scale = pb.xy_scale ? pb.xy_scale[b] : 1.0f;
      y0 = kSeeded ? 0.f : args.ref_y[b * pb.F + f];
      if (kAdjoint)
        for (int i = lane; i < S * kGenRow; i += 32) my_rows[i] = 0.f;
      acc_z = wac_z = m_s1 = m_s2 = m_n = 0.f;
      __syncwarp();
    }
    const int p_base = pb.p_begin + j * (kTraceThreads * N) + tid;
    bool has[N];
    int64_t o[N];
    V x, y, z, cx, cy;
#pragma unroll
    for (int l = 0; l < N; ++l) {
      const int p = p_base + l * kTraceThreads;
      has[l] = p < pb.p_end;
      const int q = has[l] ? p : min(p_base, pb.p_end - 1);
      o[l] = (((int64_t)b * pb.F + f) * pb.P + q) * pb.W + w;
      lane_set(x, l, __fmul_rn(pb.x.ptr[offset_of(pb.x, b, f, q, w)], xy_scale));
      lane_set(y, l, __fmul_rn(pb.y.ptr[offset_of(pb.y, b, f, q, w)], xy_scale));
      lane_set(z, l, pb.z.ptr[offset_of(pb.z, b, f, q, w)]);
      lane_set(cx, l, pb.cx.ptr[offset_of(pb.cx, b, f, q, w)]);
      lane_set(cy, l, pb.cy.ptr[offset_of(pb.cy, b, f, q, w)]);
    }
    // NOTE: no early `continue` for threads past the end of the row -- every lane of the warp
    // takes part in the shuffles below; such threads simply carry dead lanes.
    TracedGenN<V> tr = trace_guarded_gen<kAdjoint, V>(x, y, z, cx, cy, tab, S, allow_backward, pb.arith,
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
    wgt = (tr.y - V(y0)) * alive;
#pragma unroll
    for (int l = 0; l < N; ++l)
      if (!live[l]) lane_set(wgt, l, 0.f);
    m_s1 += lane_sum(wgt);
    m_s2 = lane_dot(wgt, wgt, m_s2);
    m_n += lane_sum(alive);
    if (kAdjoint) {
      // the sweep is executed by whole warps (shuffles inside): a warp skips it only if none of
      // its lanes has a live ray
      const bool warp_live = __any_sync(0xffffffffu, any_live);
      if (kSeeded && !warp_live) {
#pragma unroll
        for (int l = 0; l < N; ++l) {
          if (!has[l]) continue;
          if (args.grads.gx) args.grads.gx[o[l]] = 0.f;
          if (args.grads.gy) args.grads.gy[o[l]] = 0.f;
          if (args.grads.gz) args.grads.gz[o[l]] = 0.f;
          if (args.grads.gcx) args.grads.gcx[o[l]] = 0.f;
          if (args.grads.gcy) args.grads.gcy[o[l]] = 0.f;
        }
      }
      if (warp_live) {
        if (!any_live) {
          // this thread has no live ray: run the sweep on a harmless axial ray, seeded with 0
          for (int i = 0; i < S * 4; ++i) state[(size_t)i * stride] = V(0.f);
          tr.pre = Ray<V>{V(0.f), V(0.f), V(-tab.t_last), V(0.f), V(0.f), V(1.f)};
          tr.x = V(0.f);
          tr.y = V(0.f);
          z = V(0.f);
        } else if (!all_ok) {
          mirror_live_lane<V>(state, stride, S, tr.ok, tr.pre, z, tr.x, tr.y);
        }
        V sx(0.f), sy = alive, scx(0.f), scy(0.f);
        if (kSeeded) {
          sy = V(0.f);
#pragma unroll
          for (int l = 0; l < N; ++l) {
            if (!live[l]) continue;
            if (args.seeds.gx) lane_set(sx, l, args.seeds.gx[o[l]]);
            if (args.seeds.gy) lane_set(sy, l, args.seeds.gy[o[l]]);
            if (args.seeds.gcx) lane_set(scx, l, args.seeds.gcx[o[l]]);
            if (args.seeds.gcy) lane_set(scy, l, args.seeds.gcy[o[l]]);
          }
          wgt = alive;                        // weights: the seeds already carry everything
        }
        Sweep<V> sw = sweep_begin(tr.pre, tr.x, tr.y, sx, sy, scx, scy);
        // seed on the optical path length (row A10; trace_core_asph.cuh: OplSeed)
        OplSeed<V> path{V(0.f), V(0.f)};
        OplSeed<V> *path_seed = nullptr;
        if (kSeeded && args.seeds.gopl) {
#pragma unroll
          for (int l = 0; l < N; ++l)
            if (live[l]) lane_set(path.q, l, args.seeds.gopl[o[l]]);
          sweep_begin_opl(sw, tr.pre, path.q, V(tab.index_image));
          path_seed = &path;
        }
#pragma unroll 1
        for (int k = S - 1; k >= 0; --k) {
          const V *slot = state + (size_t)k * 4 * stride;
          const GenStep st = gen_step(tab, k);
          const AsphGrad<V> g = sweep_asphere(sw, slot[0], slot[stride], slot[2 * stride],
                                              slot[3 * stride], st.s, path_seed, V(st.index), V(st.index_next));
          float v[kGenPad];
#pragma unroll
          for (int q = 0; q < kAsphParams; ++q) {
            v[q] = lane_dot(wgt, g.p[q], 0.f);
            v[kGenPar + q] = lane_sum(g.p[q]);
          }
          v[kAsphParams] = lane_dot(wgt, g.t, 0.f);
          v[kAsphParams + 1] = lane_dot(wgt, g.mu, 0.f);
          v[kGenPar + kAsphParams] = lane_sum(g.t);
          v[kGenPar + kAsphParams + 1] = lane_sum(g.mu);
#pragma unroll
          for (int q = kGenSlots; q < kGenPad; ++q) v[q] = 0.f;
          const float total_of_lane = warp_transpose_sum24(v, lane);
          my_rows[k * kGenRow + lane] += total_of_lane;
        }
        V vx, vy, vz, vcx, vcy;
        sweep_end(sw, z, vx, vy, vz, vcx, vcy);
        acc_z += lane_sum(vz);
        wac_z = lane_dot(wgt, vz, wac_z);
        if (kSeeded) {
#pragma unroll
          for (int l = 0; l < N; ++l) {
            if (!has[l]) continue;
            const float keep = live[l] ? 1.0f : 0.0f;    // dead rays have zero gradients
            if (args.grads.gx) args.grads.gx[o[l]] = keep * lane_get(vx, l) * xy_scale;
            if (args.grads.gy) args.grads.gy[o[l]] = keep * lane_get(vy, l) * xy_scale;
            if (args.grads.gz) args.grads.gz[o[l]] = keep * lane_get(vz, l);
            if (args.grads.gcx) args.grads.gcx[o[l]] = keep * lane_get(vcx, l);
            if (args.grads.gcy) args.grads.gcy[o[l]] = keep * lane_get(vcy, l);
          }
        }
      }
    }
  }
  if (row >= 0) flush();
}

// moments[b,f,w][22 S + 5] -> rms, rms_field, gradients of the general-surface spot pass
__global__ void k_spot_finalize_gen(const double *mom_global, const float *ref_y, int B, int F, int W,
                                    int S, double n_rays, int want_grad, TlSpotOut out) {
  extern __shared__ double sh[];
  double *alpha = sh, *shift = sh + F, *rmsf = sh + 2 * F;
  const int b = blockIdx.x;
  const int n_acc = want_grad ? kGenSlots * S + 5 : 3;
  const int m0 = want_grad ? kGenSlots * S + 2 : 0;
  const double *mom = mom_global + (int64_t)b * F * W * n_acc;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    double s1 = 0.0, s2 = 0.0, n_ok = 0.0;
    for (int w = 0; w < W; ++w) {
      const double *row = mom + ((int64_t)f * W + w) * n_acc + m0;
      s1 += row[0];
      s2 += row[1];
      n_ok += row[2];
    }
    const double y0 = (double)ref_y[b * F + f];
    const double mean_rel = (s1 - (n_rays - n_ok) * y0) / n_rays;
    double ss = s2 - 2.0 * mean_rel * s1 + n_ok * mean_rel * mean_rel;
    if (ss < 0.0) ss = 0.0;
    const double rms = sqrt(ss / n_rays);
    rmsf[f] = rms;
    alpha[f] = 1.0 / ((double)F * n_rays * rms);
    shift[f] = mean_rel + (s1 - n_ok * mean_rel) / n_rays;
    out.rms_field[b * F + f] = (float)rms;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int f = 0; f < F; ++f) s += rmsf[f];
    out.rms[b] = (float)(s / F);
  }
  if (!want_grad) return;
  // outputs: per surface c, k, a[7], t summed over (f, w); mu per wavelength; z
  const int n_out = S * (kGenPar - 1) + W * S + 1;
  for (int j = threadIdx.x; j < n_out; j += blockDim.x) {
    double s = 0.0;
    if (j < S * (kGenPar - 1)) {
      const int k = j / (kGenPar - 1), q = j % (kGenPar - 1);     // q: 0 c, 1 k, 2..8 a, 9 t
      for (int f = 0; f < F; ++f) {
        double a = 0.0, bsum = 0.0;
        for (int w = 0; w < W; ++w) {
          const double *row = mom + ((int64_t)f * W + w) * n_acc + k * kGenSlots;
          a += row[q];
          bsum += row[kGenPar + q];
        }
        s += alpha[f] * (a - shift[f] * bsum);
      }
      const int64_t i = (int64_t)b * S + k;
      if (q == 0) out.gc[i] = (float)s;
      else if (q == 1) out.gk[i] = (float)s;
      else if (q == kGenPar - 2) out.gt[i] = (float)s;
      else out.ga[i * kAsphCoefs + (q - 2)] = (float)s;
    } else if (j < S * (kGenPar - 1) + W * S) {
      const int jj = j - S * (kGenPar - 1), w = jj / S, k = jj % S;
      for (int f = 0; f < F; ++f) {
        const double *row = mom + ((int64_t)f * W + w) * n_acc + k * kGenSlots;
        s += alpha[f] * (row[kGenPar - 1] - shift[f] * row[2 * kGenPar - 1]);
      }
      out.gmu[((int64_t)b * W + w) * S + k] = (float)s;
    } else {
      for (int f = 0; f < F; ++f) {
        double a = 0.0, bsum = 0.0;
        for (int w = 0; w < W; ++w) {
          const double *row = mom + ((int64_t)f * W + w) * n_acc + kGenSlots * S;
          a += row[0];
          bsum += row[1];
        }
        s += alpha[f] * (a - shift[f] * bsum);
      }
      out.gz[b] = (float)s;
    }
  }
}

// rows[b,f,w][22 S + 5] (weighted half) -> gc, gk, ga, gt [b,S(,7)], gmu [b,w,S], gz_sum [b]
__global__ void k_bwd_finalize_gen(const double *rows, TlGrads g, float *gk, float *ga, int B, int F,
                                   int W, int S) {
  const int n_acc = kGenSlots * S + 5;
  const int per_lens = S * (kGenPar - 1) + W * S + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * per_lens) return;
  const int b = i / per_lens;
  int j = i % per_lens;
  double s = 0.0;
  if (j < S * (kGenPar - 1)) {
    const int k = j / (kGenPar - 1), q = j % (kGenPar - 1);
    for (int f = 0; f < F; ++f)
      for (int w = 0; w < W; ++w) s += rows[(((int64_t)b * F + f) * W + w) * n_acc + k * kGenSlots + q];
    const int64_t o = (int64_t)b * S + k;
    if (q == 0) g.gc[o] = (float)s;
    else if (q == 1) gk[o] = (float)s;
    else if (q == kGenPar - 2) g.gt[o] = (float)s;
    else ga[o * kAsphCoefs + (q - 2)] = (float)s;
  } else if (j < S * (kGenPar - 1) + W * S) {
    j -= S * (kGenPar - 1);
    const int w = j / S, k = j % S;
    for (int f = 0; f < F; ++f)
      s += rows[(((int64_t)b * F + f) * W + w) * n_acc + k * kGenSlots + (kGenPar - 1)];
    g.gmu[((int64_t)b * W + w) * S + k] = (float)s;
  } else {
    for (int f = 0; f < F; ++f)
      for (int w = 0; w < W; ++w) s += rows[(((int64_t)b * F + f) * W + w) * n_acc + kGenSlots * S];
    g.gz_sum[b] = (float)s;
  }
}
