// k_spot_rev: the fused spot pass (trace_skew -> compute_rms2d -> .backward(), rtl:594-702) in the
// REVERSIBLE formulation of trace_core.cuh, four rays per thread (two packed fp32 pairs).
// Included by trace_kernels.cu inside its anonymous namespace.
//
// What differs from k_trace_adj<.., MODE_SPOT_GRAD, f4> (round 1), and why:
//
// * Parked state is 3 floats per ray-surface event (marching distance, cos theta, cos theta')
//   instead of 4, and the sweep rebuilds hit points and directions by walking the ray backwards
//   (sweep_sphere_rev): 2 MUFU + ~61 FMA-pipe operations per event instead of 5 + ~72.
// * A WARP, not a CTA, owns a contiguous slice of the (row, 128-ray group) list: no __syncthreads in
//   the main loop, its own surface table and parked states in shared memory, its own fp64 partial
//   row per row segment (k_reduce_rows sums them in a fixed order).
// * ACC_TMEM: the 6 S per-thread gradient accumulators do not live in registers but in TENSOR
//   MEMORY (tcgen05.st / tcgen05.ld, shape 32x32b: lane = thread, column = accumulator), which this
//   path -- it has no dense contraction, so the tensor cores idle -- otherwise leaves empty: 256 KB
//   per SM of thread-private scratch with a 12-cycle read.  That takes ~75 registers off every
//   thread, lets the adjoint sweep be a ROLLED loop (accumulator indices need not be compile-time
//   constants any more: ~350 instead of ~2 750 instructions, no instruction-cache misses) and
//   raises residency from 8 to 12 warps per SM, each still two independent dependency chains.
//   ACC_REG keeps the accumulators in registers (unrolled sweep, 8 warps per SM) for comparison.
//
// Per-event op table of the f4 instantiation: profiles/r2_*.txt (SASS mix) and DESIGN.md section 5.

enum { ACC_REG = 0, ACC_TMEM = 1, ACC_NONE = 2, ACC_OUT = 3 };   // ACC_NONE: forward sweep only (no parking, no adjoint);
                                                                 // ACC_OUT: forward trace that WRITES its six outputs (tl_trace_fwd)

struct RevArgs {
  double *partial;      // [rows, max_owners, n_acc]: row-major by (row, rank of the owning warp within the row)
  const float *ref_y;   // [B,F]
  int groups_per_row, max_owners, n_acc;
};
struct RevOutArgs : RevArgs {      // the ACC_OUT instantiation's argument (the others keep the smaller struct)
  TlTraceOut out;                  // x, y, cx, cy, ok, backward, contiguous [B,F,P,W]
};

// Surface table of one (lens, wavelength): one 32-byte record per surface, so that a step of either
// loop reads it with two 128-bit broadcast loads from one address.
struct RevSurf {
  float c, t, mu, mu2, om2, rmu;   // om2 = 1 - mu^2, rmu = 1 / mu
  int live;                        // structure mask of this surface
  int live_prev;                   // ... of the surface in front (0 for the first): rtl:626-628
};
struct RevTable {
  RevSurf *s;
  float length;   // sum |t|
  int live_last;
};

__host__ __device__ __forceinline__ size_t rev_table_floats(int S) { return 8 * (size_t)S; }

__device__ __forceinline__ RevTable load_rev_table(float *base, const TlProblem &pb, int b, int w, int lane) {
  RevTable tab;
  const int S = pb.S;
  tab.s = reinterpret_cast<RevSurf *>(base);
  __syncwarp();                                   // the previous row's readers are done
  for (int k = lane; k < S; k += 32) {
    const float m = pb.mu[((int64_t)b * pb.W + w) * S + k];
    RevSurf r;
    r.c = pb.c[(int64_t)b * S + k];
    r.t = pb.t[(int64_t)b * S + k];
    r.mu = m;
    r.mu2 = __fmul_rn(m, m);
    r.om2 = __fsub_rn(1.0f, r.mu2);      // (individually rounded: the same number fast_surface forms per thread)
    r.rmu = 1.0f / m;
    r.live = pb.live[(int64_t)b * S + k] != 0;
    r.live_prev = k > 0 && pb.live[(int64_t)b * S + k - 1] != 0;
    tab.s[k] = r;
  }
  __syncwarp();
  float len = 0.f;
  for (int k = 0; k < S; ++k) len += fabsf(tab.s[k].t);
  tab.length = len;
  tab.live_last = tab.s[S - 1].live;
  return tab;
}

// ---- tensor memory as thread-private scratch (shape 32x32b: thread i of the warp <-> TMEM lane
// 32 (warp % 4) + i; consecutive registers <-> consecutive columns) --------------------------------
__device__ __forceinline__ void tm_st4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
               "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d))
               : "memory");
}
__device__ __forceinline__ void tm_st2(uint32_t addr, float a, float b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(__float_as_uint(a)),
               "r"(__float_as_uint(b))
               : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t addr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void tm_ld2(uint32_t addr, uint32_t (&v)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t addr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(addr)
               : "memory");
}
// The loaded registers are valid only behind the wait; they are tied to it as in/out operands so that
// no use of them can be scheduled in front of it.
__device__ __forceinline__ void tm_wait_ld6(uint32_t (&a)[4], uint32_t (&b)[2]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(b[0]), "+r"(b[1])::"memory");
}
__device__ __forceinline__ void tm_wait_ld8(uint32_t (&a)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])::"memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Parked state of a thread's four rays: slot (k, comp) holds two f2 halves (rays 0-1, rays 2-3),
// each a [32]-array over the lanes of the warp: half h of slot (k, comp) of lane i is
// state[((k * NCOMP + comp) * 2 + h) * 32 + i] (conflict-free 64-bit accesses, and the two halves are
// the register pairs the packed math works on: no repacking moves).  In floats, ray l of the thread:
__device__ __forceinline__ size_t rev_ray_offset(int ncomp, int k, int comp, int l) {
  return ((size_t)(k * ncomp + comp) * 2 + (l >> 1)) * 64 + (l & 1);
}

// Exact-policy trace of one ray (rtl:594-675 statement by statement) that parks this ray's values;
// `mine` = the thread's base in the parked state, in floats.
__device__ __noinline__ Traced trace_exact_rev(float x, float y, float z, float cx, float cy, const RevTable &tab,
                                               int S, bool allow_backward, float *mine, int ncomp, int l) {
  Ray<float> r{x, y, z, cx, cy, exact_cz0(cx, cy)};
  bool ok = true, backward = false;
  for (int k = 0; k < S; ++k) {
    const RevSurf sf = tab.s[k];
    const Surface s{sf.c, sf.t, sf.mu};
    Parked<float> pk;
    exact_surface_t<false>(r, s, sf.live_prev != 0, allow_backward, ok, backward, nullptr, &pk);
    if (ncomp > 0) {                       // (ncomp == 0: the forward-only variant parks nothing)
      mine[rev_ray_offset(ncomp, k, 0, l)] = pk.dist;
      mine[rev_ray_offset(ncomp, k, 1, l)] = pk.ci;
      if (ncomp > 2) mine[rev_ray_offset(ncomp, k, 2, l)] = pk.co;
    }
  }
  Traced out;
  out.pre = r;
  exact_image(r, tab.live_last != 0, allow_backward, ok, backward);
  out.x = r.x;
  out.y = r.y;
  out.ok = ok;
  out.backward = backward;
  return out;
}

// A failed ray's parked values may hold anything: before the packed sweep runs over a thread with
// some dead lanes, every dead lane is given a live lane's states (its adjoint then stays finite and,
// seeded with 0, contributes exact zeros).
__device__ __noinline__ void mirror_live_lane_rev(float *mine, int S, int ncomp, const bool *ok, Ray<f4> &pre,
                                                  f4 &x_img, f4 &y_img) {
  int src = 0;
  for (int l = 0; l < 4; ++l)
    if (ok[l]) src = l;
  for (int k = 0; k < S; ++k)
    for (int comp = 0; comp < ncomp; ++comp) {
      const float v = mine[rev_ray_offset(ncomp, k, comp, src)];
      for (int l = 0; l < 4; ++l)
        if (!ok[l]) mine[rev_ray_offset(ncomp, k, comp, l)] = v;
    }
  f4 *comp[8] = {&pre.x, &pre.y, &pre.z, &pre.cx, &pre.cy, &pre.cz, &x_img, &y_img};
  for (int j = 0; j < 8; ++j) {
    const float v = lane_get(*comp[j], src);
    for (int l = 0; l < 4; ++l)
      if (!ok[l]) lane_set(*comp[j], l, v);
  }
}

// Everything about the rays of one (lens, field, wavelength) row that does not depend on the pupil
// index, gathered once per row: the generic strided load (five tensors x four rays x four 64-bit
// index products, plus the vignetting / aiming records) cost ~350 instructions per 128-ray group,
// 6 % of the kernel (profiles/r2c).  With it a ray costs two indexed loads and the few map steps.
struct RowRays {
  const float *x, *y;                    // element (b, f, 0, w) of the pupil coordinates
  float z0, cx0, cy0, cz0;               // z, cx, cy (and fast_cz0 of them) when all three are pupil-invariant
  float vig[3], aim[3], xy_scale;        // apply_vignetting record, ray-aiming record (load_pupil_point)
};
// (strides and the presence of the optional tables are kernel parameters: read from the constant bank
// where they are used, they cost no registers)
__device__ __forceinline__ bool row_uniform_zc(const TlProblem &pb) {
  return pb.z.stride[2] == 0 && pb.cx.stride[2] == 0 && pb.cy.stride[2] == 0;
}

__device__ __forceinline__ RowRays load_row_rays(const TlProblem &pb, int b, int f, int w) {
  RowRays r;
  r.x = pb.x.ptr + offset_of(pb.x, b, f, 0, w);
  r.y = pb.y.ptr + offset_of(pb.y, b, f, 0, w);
  r.z0 = pb.z.ptr[offset_of(pb.z, b, f, 0, w)];
  r.cx0 = pb.cx.ptr[offset_of(pb.cx, b, f, 0, w)];
  r.cy0 = pb.cy.ptr[offset_of(pb.cy, b, f, 0, w)];
  r.cz0 = fast_cz0(r.cx0, r.cy0);
  r.xy_scale = pb.xy_scale ? pb.xy_scale[b] : 1.0f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    r.vig[j] = pb.vig ? pb.vig[((int64_t)b * pb.F + f) * 3 + j] : 0.f;
    r.aim[j] = pb.aim ? pb.aim[(((int64_t)b * pb.F + f) * pb.W + w) * 3 + j] : 0.f;
  }
  return r;
}

// load_pupil_point (trace_kernels.cu) on a row record: the same individually rounded steps
__device__ __forceinline__ void row_pupil_point(const TlProblem &pb, const RowRays &r, int q, float &x, float &y) {
  x = r.x[q * pb.x.stride[2]];
  y = r.y[q * pb.y.stride[2]];
  if (pb.vig) {
    x = __fmul_rn(x, r.vig[0]);
    y = __fadd_rn(__fmul_rn(y, r.vig[1]), r.vig[2]);
  }
  if (pb.aim) {
    x = fminf(fmaxf(__fmul_rn(x, r.aim[0]), -2.0f), 2.0f);
    y = fminf(fmaxf(__fadd_rn(__fmul_rn(y, r.aim[1]), r.aim[2]), -2.0f), 2.0f);
  }
  if (pb.xy_scale) {
    x = __fmul_rn(x, r.xy_scale);
    y = __fmul_rn(y, r.xy_scale);
  }
}

template <int NS_MAX, int NW, int ACC, int NCOMP, class ARGS = RevArgs>
__global__ void __launch_bounds__(NW * 32, 1)
k_spot_rev(TlProblem pb, ARGS args) {
  extern __shared__ __align__(16) float smem_rev[];
  __shared__ uint32_t tmem_slot;
  using V = f4;
  constexpr int N = 4;
  constexpr int NA = (ACC == ACC_REG) ? NS_MAX : 1;
  constexpr bool GRAD = ACC == ACC_REG || ACC == ACC_TMEM;      // ACC_NONE: trace + spot moments only (tl_spot_accumulate
                                                                // without gradients); ACC_OUT: trace_skew's forward, materialised
  constexpr bool WRITE = ACC == ACC_OUT;
  static_assert(GRAD ? NCOMP >= 2 : NCOMP == 0, "parked components: 2 or 3 with the adjoint, none without");
  constexpr uint32_t kTmemCols = 512;
  const int S = pb.S;
  const int lane = threadIdx.x & 31;
  // (a broadcast from lane 0: the compiler then knows the warp index -- and every table / TMEM address
  // derived from it -- is warp-uniform and keeps them in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const bool allow_backward = pb.allow_backward_rays != 0;
  const int n_acc = args.n_acc;

  uint32_t tacc = 0;      // this warp's accumulator columns: lane quarter warp % 4, column block warp / 4
  if constexpr (ACC == ACC_TMEM) {
    if (warp == 0) {
      const uint32_t slot = (uint32_t)__cvta_generic_to_shared(&tmem_slot);
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tacc = tmem_slot + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * (kTmemCols / ((NW + 3) / 4));
  }

  const size_t tab_floats = rev_table_floats(S);
  const size_t per_warp = tab_floats + (size_t)S * NCOMP * 2 * 64;     // (NCOMP == 0: the table alone)
  float *wbase = smem_rev + warp * per_warp;
  f2 *state = reinterpret_cast<f2 *>(wbase + tab_floats) + lane;     // half h of slot (k, comp): state[((k NCOMP + comp) 2 + h) 32]
  float *mine = wbase + tab_floats + 2 * lane;                       // the same, in floats (rev_ray_offset)

  const int64_t total = (int64_t)pb.B * pb.F * pb.W * args.groups_per_row;
  const int64_t n_warps = (int64_t)gridDim.x * NW;
  const int64_t wg = (int64_t)blockIdx.x * NW + warp;
  const int64_t g_begin = total * wg / n_warps;
  const int64_t g_end = total * (wg + 1) / n_warps;

  // register accumulators (ACC_REG; every index is a compile-time constant after unrolling)
  float acc[6][NA];      // wac_c, acc_c, wac_t, acc_t, wac_mu, acc_mu
  float acc_z = 0.f, wac_z = 0.f, m_s1 = 0.f, m_s2 = 0.f, m_n = 0.f;
  RevTable tab;
  RowRays rr;
  float y0 = 0.f;
  int row = -1, b = 0, f = 0, w = 0;

  auto reset = [&]() {
    if constexpr (ACC == ACC_REG) {
#pragma unroll
      for (int k = 0; k < NA; ++k)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[j][k] = 0.f;
    } else if constexpr (ACC == ACC_TMEM) {
      for (int k = 0; k < S; ++k) {
        tm_st4(tacc + 6 * k, 0.f, 0.f, 0.f, 0.f);
        tm_st2(tacc + 6 * k + 4, 0.f, 0.f);
      }
      tm_wait_st();
    }
    acc_z = wac_z = m_s1 = m_s2 = m_n = 0.f;
  };

  // the warp's sums of this row segment -> one fp64 partial row
  auto flush = [&]() {
    if constexpr (WRITE) return;      // (the output-writing variant keeps no sums)
    // partial row of (row, this warp's rank among the warps that own a piece of the row): the reducer
    // then sums a row's partials from consecutive memory, without recomputing any slice bounds
    const int64_t first_owner = owner_of((int64_t)row * args.groups_per_row, total, n_warps);
    double *dst = args.partial + ((int64_t)row * args.max_owners + (wg - first_owner)) * n_acc;
    if constexpr (ACC == ACC_REG) {
      constexpr int kPerBatch = 5;             // 6 values per surface, five surfaces per transpose
#pragma unroll
      for (int i = 0; i < (NS_MAX + kPerBatch - 1) / kPerBatch; ++i) {
        float v[32];
#pragma unroll
        for (int q = 0; q < kPerBatch; ++q) {
          const int k = i * kPerBatch + q;
#pragma unroll
          for (int j = 0; j < 6; ++j) v[6 * q + j] = k < NA ? acc[j][k < NA ? k : 0] : 0.f;
        }
        v[30] = v[31] = 0.f;
        const float sum = warp_transpose_sum(v, lane);
        const int k = i * kPerBatch + lane / 6, j = lane % 6;
        if (lane < 30 && k < S) dst[j * S + k] = (double)sum;
      }
    } else if constexpr (ACC == ACC_TMEM) {
      tm_wait_st();
      for (int c0 = 0; c0 < 6 * S; c0 += 32) {      // (columns behind 6 S hold stale values: discarded below)
        float v[32];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t u[8];
          tm_ld8(tacc + c0 + 8 * q, u);
          tm_wait_ld8(u);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[8 * q + j] = __uint_as_float(u[j]);
        }
        const float sum = warp_transpose_sum(v, lane);
        const int i = c0 + lane;
        if (i < 6 * S) dst[(i % 6) * S + i / 6] = (double)sum;
      }
    }
    const float t2 = warp_sum(m_s1), t3 = warp_sum(m_s2), t4 = warp_sum(m_n);
    if constexpr (GRAD) {
      const float t0 = warp_sum(wac_z), t1 = warp_sum(acc_z);
      if (lane == 0) {
        dst[6 * S] = (double)t0;
        dst[6 * S + 1] = (double)t1;
        dst[6 * S + 2] = (double)t2;
        dst[6 * S + 3] = (double)t3;
        dst[6 * S + 4] = (double)t4;
      }
    } else if (lane == 0) {                  // MODE_SPOT_EVAL's three moments
      dst[0] = (double)t2;
      dst[1] = (double)t3;
      dst[2] = (double)t4;
    }
  };

  // (row, group within the row) of the slice's first item by one division; then counted up -- the 64-bit
  // division per group was ~30 instructions of every 128-ray group
  int r = (int)(g_begin / args.groups_per_row), j = (int)(g_begin % args.groups_per_row) - 1;
  for (int64_t g = g_begin; g < g_end; ++g) {
    if (++j == args.groups_per_row) {
      j = 0;
      ++r;
    }
    if (r != row) {
      if (row >= 0) flush();
      row = r;
      w = r % pb.W;
      f = (r / pb.W) % pb.F;
      b = r / (pb.W * pb.F);
      tab = load_rev_table(wbase, pb, b, w, lane);
      rr = load_row_rays(pb, b, f, w);
      y0 = WRITE ? 0.f : args.ref_y[b * pb.F + f];
      reset();
    }
    // lane l of this thread = pupil point p_base + 32 l; past the end of the slice: a copy of its
    // last point, traced but seeded with 0 (every thread runs every step: the TMEM accesses are
    // warp-collective)
    const int p_base = pb.p_begin + j * (32 * N) + lane;
    bool has[N];
    V x, y, z, cx, cy;
#pragma unroll
    for (int l = 0; l < N; ++l) {
      const int p = p_base + l * 32;
      has[l] = p < pb.p_end;
      const int q = has[l] ? p : pb.p_end - 1;
      float px, py;
      row_pupil_point(pb, rr, q, px, py);
      lane_set(x, l, px);
      lane_set(y, l, py);
      if (!row_uniform_zc(pb)) {
        lane_set(z, l, pb.z.ptr[offset_of(pb.z, b, f, q, w)]);
        lane_set(cx, l, pb.cx.ptr[offset_of(pb.cx, b, f, q, w)]);
        lane_set(cy, l, pb.cy.ptr[offset_of(pb.cy, b, f, q, w)]);
      }
    }
    if (row_uniform_zc(pb)) {
      z = V(rr.z0);
      cx = V(rr.cx0);
      cy = V(rr.cy0);
    }

    // ---- forward: fast policy for all four lanes, parking (dist, cos[, cos']) per surface ----
    Ray<V> pre;
    V x_img, y_img;
    bool ok[N], clear[N];
#pragma unroll
    for (int l = 0; l < N; ++l) clear[l] = false;
    if (pb.arith == TL_ARITH_GUARDED) {
      Ray<V> ray{x, y, z, cx, cy, row_uniform_zc(pb) ? V(rr.cz0) : fast_cz0(cx, cy)};
      V min_cos2(1.0f), min_cz(1.0f), min_travel(3.0e38f);
      const float4 *rec = reinterpret_cast<const float4 *>(tab.s);                // running pointers: no index
      f2 *slot = state;                                                           // multiplications in the loop
      for (int k = 0; k < S; ++k, rec += 2, slot += NCOMP * 64) {
        const float4 s0 = rec[0];           // c, t, mu, mu2
        const float4 s1 = rec[1];           // om2, rmu, live, live_prev
        V travel;
        Parked<V> pk;
        fast_surface_rev(ray, V(s0.x), V(s0.z), V(s0.w), V(s1.x), V(s0.y), min_cos2, min_cz, travel, pk);
        if constexpr (GRAD) {
          slot[0] = pk.dist.a;
          slot[32] = pk.dist.b;
          slot[64] = pk.ci.a;
          slot[96] = pk.ci.b;
        }
        if constexpr (NCOMP > 2) {
          slot[128] = pk.co.a;
          slot[160] = pk.co.b;
        }
        if (__float_as_int(s1.w) != 0) min_travel = fmin2(min_travel, travel);
      }
      pre = ray;
      const V travel = fast_image(ray);
      if (tab.live_last) min_travel = fmin2(min_travel, travel);
      x_img = ray.x;
      y_img = ray.y;
      const V probe = (ray.x + ray.y) + (ray.cx + ray.cy);
#pragma unroll
      for (int l = 0; l < N; ++l) {
        const float band = kBandTravelRel * fmaxf(1.0f, tab.length + fabsf(lane_get(z, l)));
        clear[l] = (lane_get(min_cos2, l) > kGuard + kBandCos2) && (lane_get(min_cz, l) > kBandCz) &&
                   (lane_get(min_travel, l) > band) && (fabsf(lane_get(probe, l)) < 3.0e38f);
      }
    }
    bool any_live = false, all_ok = true, any_ok = false;
    [[maybe_unused]] bool bw[WRITE ? N : 1];      // ray_backward (ACC_OUT): false on the fast path by construction of its guard band
    V alive;
#pragma unroll
    for (int l = 0; l < N; ++l) {
      ok[l] = true;
      if constexpr (WRITE) bw[l] = false;
      if (!clear[l]) {      // not clearly good everywhere: this lane alone, exact policy
        const Traced one = trace_exact_rev(lane_get(x, l), lane_get(y, l), lane_get(z, l), lane_get(cx, l),
                                           lane_get(cy, l), tab, S, allow_backward, mine, NCOMP, l);
        lane_set(pre.x, l, one.pre.x);
        lane_set(pre.y, l, one.pre.y);
        lane_set(pre.z, l, one.pre.z);
        lane_set(pre.cx, l, one.pre.cx);
        lane_set(pre.cy, l, one.pre.cy);
        lane_set(pre.cz, l, one.pre.cz);
        lane_set(x_img, l, one.x);
        lane_set(y_img, l, one.y);
        ok[l] = one.ok;
        if constexpr (WRITE) bw[l] = one.backward;
      }
      const bool live = ok[l] && has[l];
      any_live = any_live || live;
      all_ok = all_ok && ok[l];
      any_ok = any_ok || ok[l];
      lane_set(alive, l, live ? 1.0f : 0.0f);
    }
    if constexpr (WRITE) {                    // trace_skew's six outputs (rtl:672-675), contiguous [B,F,P,W]
#pragma unroll
      for (int l = 0; l < N; ++l) {
        if (!has[l]) continue;
        const int64_t o = (((int64_t)b * pb.F + f) * pb.P + (p_base + l * 32)) * pb.W + w;
        args.out.x[o] = lane_get(x_img, l);
        args.out.y[o] = lane_get(y_img, l);
        args.out.cx[o] = lane_get(pre.cx, l);
        args.out.cy[o] = lane_get(pre.cy, l);
        args.out.ok[o] = ok[l];
        args.out.backward[o] = bw[l];
      }
      continue;
    }
    V wgt = (y_img - V(y0)) * alive;
#pragma unroll
    for (int l = 0; l < N; ++l)
      if (lane_get(alive, l) == 0.0f) lane_set(wgt, l, 0.f);      // a dead lane's y may be anything
    m_s1 += lane_sum(wgt);
    m_s2 = lane_dot(wgt, wgt, m_s2);
    m_n += lane_sum(alive);
    if constexpr (GRAD) {                     // (forward sweep only: the row's three moments are all there is)
    if (!all_ok && any_ok) {
      // (copies: the helper takes addresses, and address-taken variables live in local memory -- of `pre`,
      // `x_img`, `y_img` and `ok` themselves that cost 8 STL.128 + 8 LDL per group on the common path)
      bool ok_copy[N];
#pragma unroll
      for (int l = 0; l < N; ++l) ok_copy[l] = ok[l];
      Ray<V> pre_copy = pre;
      V x_copy = x_img, y_copy = y_img;
      mirror_live_lane_rev(mine, S, NCOMP, ok_copy, pre_copy, x_copy, y_copy);
      pre = pre_copy;
      x_img = x_copy;
      y_img = y_copy;
    }

    // ---- adjoint: unit seed on y, backward walk ----
    // (a thread none of whose rays is alive still runs the sweep -- the accumulator accesses are
    // collective -- but adds nothing: its state may hold non-finite values)
    SweepRev<V> sw = sweep_begin_rev(pre, x_img, y_img, V(0.f), alive, V(0.f), V(0.f));
    auto step = [&](const float4 *rec, const f2 *slot) {
      const float4 s0 = rec[0];
      const float4 s1 = rec[1];
      const V dist(slot[0], slot[32]), ci(slot[64], slot[96]);
      if constexpr (NCOMP > 2) {
        const V co(slot[128], slot[160]);
        return sweep_sphere_rev(sw, Parked<V>{dist, ci, co}, V(s0.x), V(s0.y), V(s0.z), V(s1.y));
      } else {
        return sweep_sphere_rev2(sw, dist, ci, V(s0.x), V(s0.y), V(s0.z), V(s0.w), V(s1.x), V(s1.y));
      }
    };
    if constexpr (ACC == ACC_REG) {
#pragma unroll
      for (int k = NS_MAX - 1; k >= 0; --k) {
        if (k >= S) continue;
        const SurfaceGrad<V> gr = step(reinterpret_cast<const float4 *>(tab.s) + 2 * k, state + (size_t)k * NCOMP * 64);
        if (any_live) {
          acc[0][k] = lane_dot(wgt, gr.c, acc[0][k]);
          acc[1][k] += lane_sum(gr.c);
          acc[2][k] = lane_dot(wgt, gr.t, acc[2][k]);
          acc[3][k] += lane_sum(gr.t);
          acc[4][k] = lane_dot(wgt, gr.mu, acc[4][k]);
          acc[5][k] += lane_sum(gr.mu);
        }
      }
    } else if constexpr (ACC == ACC_TMEM) {
      const float4 *rec = reinterpret_cast<const float4 *>(tab.s) + 2 * (S - 1);
      const f2 *slot = state + (size_t)(S - 1) * NCOMP * 64;
      uint32_t tcol = tacc + 6 * (S - 1);
      for (int k = S - 1; k >= 0; --k, rec -= 2, slot -= NCOMP * 64, tcol -= 6) {
        uint32_t a4[4], a2[2];
        tm_ld4(tcol, a4);
        tm_ld2(tcol + 4, a2);
        const SurfaceGrad<V> gr = step(rec, slot);
        tm_wait_ld6(a4, a2);
        float s0 = __uint_as_float(a4[0]), s1 = __uint_as_float(a4[1]), s2 = __uint_as_float(a4[2]),
              s3 = __uint_as_float(a4[3]), s4 = __uint_as_float(a2[0]), s5 = __uint_as_float(a2[1]);
        if (any_live) {
          s0 = lane_dot(wgt, gr.c, s0);
          s1 += lane_sum(gr.c);
          s2 = lane_dot(wgt, gr.t, s2);
          s3 += lane_sum(gr.t);
          s4 = lane_dot(wgt, gr.mu, s4);
          s5 += lane_sum(gr.mu);
        }
        tm_st4(tcol, s0, s1, s2, s3);
        tm_st2(tcol + 4, s4, s5);
      }
      tm_wait_st();
    }
    V ax, ay, az, acx, acy;
    sweep_end_rev(sw, ax, ay, az, acx, acy);
    if (any_live) {
      acc_z += lane_sum(az);
      wac_z = lane_dot(wgt, az, wac_z);
    }
    }      // GRAD
  }
  if (row >= 0) flush();

  if constexpr (ACC == ACC_TMEM) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
  }
}

// partial[row, owner rank, n_acc] -> dst[row, n_acc], fixed summation order (rank order = pupil order).
// One CTA per row, a thread per slot: a row's partials are consecutive, the loads of a warp
// coalesce and four are in flight per thread.  (The first version walked [warp, segment] records with
// two 64-bit divisions per term: 23 us of a 270 us step.)
__global__ void __launch_bounds__(128)
k_reduce_owner_rows(const double *partial, double *dst, int n_rows, int groups_per_row, int64_t n_warps, int max_owners,
                    int n_acc) {
  __shared__ double part[4][256];                  // (n_acc <= 6 * 32 + 5)
  const int row = blockIdx.x;
  const int lane = threadIdx.x & 31, quarter = threadIdx.x >> 5;
  const int64_t total = (int64_t)n_rows * groups_per_row;
  const int64_t first = owner_of((int64_t)row * groups_per_row, total, n_warps);
  const int64_t last = owner_of((int64_t)(row + 1) * groups_per_row - 1, total, n_warps);
  const int count = (int)(last - first + 1);
  const bool sparse = n_warps > total;             // some warps own nothing: their records were never written
  const double *src = partial + (int64_t)row * max_owners * n_acc;
  // warp `quarter` sums the owners quarter, quarter + 4, ... of every slot (all loads independent), then
  // the four partial sums are added in a fixed order
  for (int slot = lane; slot < n_acc; slot += 32) {
    double acc = 0.0;
    if (!sparse) {
#pragma unroll 8
      for (int q = quarter; q < count; q += 4) acc += src[(int64_t)q * n_acc + slot];      // (loads in flight together)
    } else {
      for (int q = quarter; q < count; q += 4) {
        const int64_t wgq = first + q;
        if (total * (wgq + 1) / n_warps > total * wgq / n_warps) acc += src[(int64_t)q * n_acc + slot];
      }
    }
    part[quarter][slot] = acc;
  }
  __syncthreads();
  for (int slot = threadIdx.x; slot < n_acc; slot += blockDim.x)
    dst[(int64_t)row * n_acc + slot] = (part[0][slot] + part[1][slot]) + (part[2][slot] + part[3][slot]);
}

typedef const void *RevKernelPtr;      // (TlProblem, RevArgs), or (TlProblem, RevOutArgs) for the ACC_OUT instantiation

struct RevPlan {
  RevKernelPtr kernel = nullptr;
  const char *name = "";
  int n_warps_cta = 8, n_blocks = 1, groups_per_row = 1, max_owners = 1, n_acc = 0;
  size_t smem = 0, partial_bytes = 0;
};

// Variant: warps per CTA (one persistent CTA per SM), where the accumulators live, and how many
// values are parked per event.  TL_REV picks one for experiments ("reg8", "tmem8", "tmem12",
// "tmem12c2", "tmem16"); the default is the fastest measured that fits.
struct RevVariant {
  const char *name;
  RevKernelPtr kernel;
  int nw, ncomp;
};

int plan_rev(const TlProblem &pb, RevPlan &pl, int want_grad = 1) {
  DeviceInfo info;
  int rc = device_info(info);
  if (rc) return rc;
  const int S = pb.S;
  // in the order measured on the B200 (config 2, tools/rev_variants.py: reference heights + this kernel + row
  // reduction as a graph, L2 flushed; profiles/r2_rev_variants.json): tmem12 0.2147 ms, tmem12c2 0.2188, tmem16
  // 0.2224 (128 registers: spills), tmem8 0.2263, reg8 0.2338 -- and tmem14c2 0.2467, tmem10 0.2538: warp counts
  // that are not a multiple of the four schedulers lose 15 %.  (Round 1's k_trace_adj: 0.2988.)  Residency matters
  // little (8 -> 12 warps per SM: 5 %): the pass is bound by register-operand bandwidth (tools/microbench4.cu).
  const RevVariant variants[] = {
      {"tmem12", (const void *)k_spot_rev<16, 12, ACC_TMEM, 3>, 12, 3},
      {"tmem14c2", (const void *)k_spot_rev<16, 14, ACC_TMEM, 2>, 14, 2},
      {"tmem16", (const void *)k_spot_rev<16, 16, ACC_TMEM, 2>, 16, 2},
      {"tmem12c2", (const void *)k_spot_rev<16, 12, ACC_TMEM, 2>, 12, 2},
      {"tmem10", (const void *)k_spot_rev<16, 10, ACC_TMEM, 3>, 10, 3},
      {"tmem8", (const void *)k_spot_rev<16, 8, ACC_TMEM, 3>, 8, 3},
      {"tmem4", (const void *)k_spot_rev<16, 4, ACC_TMEM, 3>, 4, 3},      // the only one whose parked state fits for 19..32 surfaces
      {"reg8", S <= 12 ? (const void *)k_spot_rev<12, 8, ACC_REG, 3> : (const void *)k_spot_rev<16, 8, ACC_REG, 3>, 8, 3},
  };
  // forward only (no parked state, 115 registers): 16 warps per SM; 20 / 24 warps measured 1-2 % slower
  // (tools/profile_forward.py: 0.0940 / 0.0954 / 0.0959 ms against 0.1261 ms for round 1's k_trace_adj<SPOT_EVAL,f4>)
  const RevVariant eval_variant = {"eval16", (const void *)k_spot_rev<16, 16, ACC_NONE, 0>, 16, 0};
  const RevVariant out_variant = {"out16", (const void *)k_spot_rev<16, 16, ACC_OUT, 0, RevOutArgs>, 16, 0};      // want_grad == 2: tl_trace_fwd
  const char *env = want_grad == 1 ? getenv("TL_REV") : nullptr;
  const RevVariant *pick = want_grad == 1 ? nullptr : (want_grad == 2 ? &out_variant : &eval_variant);
  for (const RevVariant &v : variants) {
    if (pick) break;
    const size_t smem = (size_t)v.nw * (rev_table_floats(S) + (size_t)S * v.ncomp * 2 * 64) * sizeof(float);
    const int tmem_cols = 512 / ((v.nw + 3) / 4);                               // per warp: TMEM column blocks
    const bool is_reg = !strcmp(v.name, "reg8");                                // (register accumulators: 16 surfaces)
    const bool fits = smem <= 227 * 1024 && (is_reg ? S <= 16 : 6 * S <= tmem_cols - 32);   // (flush reads whole 32-column blocks)
    if (env ? !strcmp(env, v.name) : fits) {
      if (!fits) return fail(TL_ERR_INVALID, "TL_REV variant does not fit this surface count%s");
      pick = &v;
      break;
    }
  }
  if (!pick) return fail(TL_ERR_INVALID, "no variant of the reversible spot kernel fits (or unknown TL_REV)%s");
  pl.kernel = pick->kernel;
  pl.name = pick->name;
  const int nw = pick->nw;
  pl.n_warps_cta = nw;
  pl.n_acc = n_acc_of(want_grad == 1 ? MODE_SPOT_GRAD : MODE_SPOT_EVAL, S);
  pl.smem = (size_t)nw * (rev_table_floats(S) + (size_t)S * pick->ncomp * 2 * 64) * sizeof(float);
  TL_CHECK_CUDA(cudaFuncSetAttribute((const void *)pl.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)pl.smem));
  const int group = 32 * 4;
  pl.groups_per_row = (pb.p_end - pb.p_begin + group - 1) / group;
  const int64_t total = (int64_t)pb.B * pb.F * pb.W * pl.groups_per_row;
  int64_t n_blocks = info.sms;
  const int64_t need = (total + nw - 1) / nw;
  if (n_blocks > need) n_blocks = need;
  pl.n_blocks = (int)n_blocks;
  const int64_t n_warps = n_blocks * nw;
  // warps that own a piece of one row: a row is groups_per_row consecutive items, a slice at least
  // floor(total / n_warps) of them (when that is 0, some slices are empty: the reducer checks)
  const int64_t shortest = total / n_warps;
  const int64_t owners = shortest > 0 ? (pl.groups_per_row + shortest - 1) / shortest + 1
                                      : 2 * (int64_t)pl.groups_per_row + 2;      // (ranks count the empty slices too)
  pl.max_owners = (int)(owners < n_warps ? owners : n_warps);
  pl.partial_bytes = align8((size_t)pb.B * pb.F * pb.W * pl.max_owners * pl.n_acc * sizeof(double));
  return TL_OK;
}

bool use_rev_kernel(const TlProblem &pb) { return pb.S <= TL_MAX_SURFACES_SPOT && !getenv("TL_NO_REV"); }
// the forward-only variant holds nothing per surface but the table: any surface count the forward entry points take
bool use_rev_eval_kernel(const TlProblem &pb) { return pb.S <= TL_MAX_SURFACES_FWD && !getenv("TL_NO_REV"); }

int launch_spot_rev(const TlProblem &pb, const RevPlan &pl, const float *ref_y, double *partial, double *moments,
                    cudaStream_t stream) {
  RevArgs args{};
  args.partial = partial;
  args.ref_y = ref_y;
  args.groups_per_row = pl.groups_per_row;
  args.max_owners = pl.max_owners;
  args.n_acc = pl.n_acc;
  TlProblem pb_copy = pb;
  void *params[] = {(void *)&pb_copy, (void *)&args};
  TL_CHECK_CUDA(cudaLaunchKernel((const void *)pl.kernel, dim3(pl.n_blocks), dim3(pl.n_warps_cta * 32), params,
                                 pl.smem, stream));
  g_launches++;
  const int rows = pb.B * pb.F * pb.W;
  k_reduce_owner_rows<<<rows, 128, 0, stream>>>(partial, moments, rows, pl.groups_per_row,
                                                pl.n_blocks * pl.n_warps_cta, pl.max_owners, pl.n_acc);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

// trace_skew's forward, materialised (tl_trace_fwd): the forward-only kernel writing its six outputs, whole pupil
int launch_trace_rev(const TlProblem &pb, const TlTraceOut &out, cudaStream_t stream) {
  TlProblem full = pb;
  full.p_begin = 0;
  full.p_end = pb.P;
  RevPlan pl;
  int rc = plan_rev(full, pl, 2);
  if (rc) return rc;
  RevOutArgs args{};
  args.groups_per_row = pl.groups_per_row;
  args.max_owners = pl.max_owners;
  args.n_acc = pl.n_acc;
  args.out = out;
  void *params[] = {(void *)&full, (void *)&args};
  TL_CHECK_CUDA(cudaLaunchKernel((const void *)pl.kernel, dim3(pl.n_blocks), dim3(pl.n_warps_cta * 32), params,
                                 pl.smem, stream));
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}
