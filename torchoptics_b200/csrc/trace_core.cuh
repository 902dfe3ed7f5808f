// Per-ray arithmetic of the sequential skew-ray trace: forward (two arithmetic
// policies) and the hand-derived adjoint of one surface.
//
// Everything here is a pure function of a ray's registers; the kernels in
// trace_kernels.cu decide how rays map to threads, where per-surface state is
// parked and how gradients are reduced.  The file also compiles as plain C++
// (g++ -ffp-contract=off) so that tests/hostcore can run the very same code on
// the CPU against the oracle -- a test aid only, never a product fallback.
//
// Reference behaviour followed (file: /root/reference/torchlens/ray_tracing_lite.py):
//   march      find_marching_distance_spherical   rtl:525-545
//   advance    update_ray_coordinates             rtl:514-522
//   refract    apply_snell_spherical              rtl:548-571
//   park       reset_bad_rays                     rtl:574-591
//   loop/masks trace_skew                         rtl:594-675
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TL_HD __host__ __device__ __forceinline__
#else
#define TL_HD inline
struct float2 { float x, y; };   // host test build only
#endif

namespace tl {

constexpr float kGuard = 1e-6f;   // rtl:530, rtl:552: eps of every mask predicate
// Guard bands of the GUARDED policy: a ray takes the contracted fast path only if
// every predicate quantity clears its threshold by this much (cos^2-like
// quantities are O(1); z travel is compared against the system length).
constexpr float kBandCos2 = 1e-4f;
constexpr float kBandTravelRel = 1e-5f;

// ---------------------------------------------------------------------------
// exact scalar ops: one IEEE-754 round-to-nearest operation each, never fused
// ---------------------------------------------------------------------------
TL_HD float xmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
TL_HD float xadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
TL_HD float xsub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
TL_HD float xdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
TL_HD float xsqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}

// ---------------------------------------------------------------------------
// fast ops (contracted, approximate reciprocal / rsqrt on the MUFU unit)
// ---------------------------------------------------------------------------
TL_HD float ffma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
TL_HD float frcp(float a) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#else
  return 1.0f / a;
#endif
}
TL_HD float frsqrt(float a) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#else
  return 1.0f / sqrtf(a);
#endif
}
// Two rays per thread: a pair of fp32 lanes that maps onto Blackwell's packed
// FFMA2 / FMUL2 / FADD2 (fma.rn.f32x2 ...).  The FMA pipe retires a packed
// instruction at the same flop rate as two scalar ones, but it costs ONE issue
// slot, which frees the scheduler for the MUFU / shared-memory / select traffic of
// the trace (measured: tools/microbench.cu).  Each lane is rounded exactly like the
// scalar op, so a pair gives bit-identical results to two scalar rays.
struct alignas(8) f2 {
  float2 v;
  TL_HD f2() {}
  TL_HD f2(float s) { v.x = s; v.y = s; }
  TL_HD f2(float a, float b) { v.x = a; v.y = b; }
};
TL_HD f2 operator-(f2 a) { return f2(-a.v.x, -a.v.y); }
TL_HD f2 operator+(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
  f2 r; r.v = __fadd2_rn(a.v, b.v); return r;
#else
  return f2(a.v.x + b.v.x, a.v.y + b.v.y);
#endif
}
TL_HD f2 operator-(f2 a, f2 b) { return a + (-b); }
TL_HD f2 operator*(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
  f2 r; r.v = __fmul2_rn(a.v, b.v); return r;
#else
  return f2(a.v.x * b.v.x, a.v.y * b.v.y);
#endif
}
TL_HD f2 ffma(f2 a, f2 b, f2 c) {
#if defined(__CUDA_ARCH__)
  f2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r;
#else
  return f2(fmaf(a.v.x, b.v.x, c.v.x), fmaf(a.v.y, b.v.y, c.v.y));
#endif
}
TL_HD f2 frcp(f2 a) { return f2(frcp(a.v.x), frcp(a.v.y)); }
TL_HD f2 frsqrt(f2 a) { return f2(frsqrt(a.v.x), frsqrt(a.v.y)); }
TL_HD f2 fmin2(f2 a, f2 b) { return f2(fminf(a.v.x, b.v.x), fminf(a.v.y, b.v.y)); }

// Four rays per thread: two independent pairs.  Every operation is issued once per
// pair, so a thread carries two independent dependency chains (ILP 2) -- the trace
// is a long chain of dependent FMAs and MUFUs, and this is what hides their latency
// at the low occupancy the per-thread state allows.
struct alignas(16) f4 {
  f2 a, b;
  TL_HD f4() {}
  TL_HD f4(float s) : a(s), b(s) {}
  TL_HD f4(f2 a_, f2 b_) : a(a_), b(b_) {}
};
TL_HD f4 operator-(f4 v) { return f4(-v.a, -v.b); }
TL_HD f4 operator+(f4 p, f4 q) { return f4(p.a + q.a, p.b + q.b); }
TL_HD f4 operator-(f4 p, f4 q) { return f4(p.a - q.a, p.b - q.b); }
TL_HD f4 operator*(f4 p, f4 q) { return f4(p.a * q.a, p.b * q.b); }
TL_HD f4 ffma(f4 p, f4 q, f4 r) { return f4(ffma(p.a, q.a, r.a), ffma(p.b, q.b, r.b)); }
TL_HD f4 frcp(f4 v) { return f4(frcp(v.a), frcp(v.b)); }
TL_HD f4 frsqrt(f4 v) { return f4(frsqrt(v.a), frsqrt(v.b)); }
TL_HD f4 fmin2(f4 p, f4 q) { return f4(fmin2(p.a, q.a), fmin2(p.b, q.b)); }

// lane access (indices are compile-time constants after unrolling)
template <class V> struct LaneCount;
template <> struct LaneCount<float> { static constexpr int value = 1; };
template <> struct LaneCount<f2> { static constexpr int value = 2; };
template <> struct LaneCount<f4> { static constexpr int value = 4; };
TL_HD float lane_get(const float &v, int) { return v; }
TL_HD float lane_get(const f2 &v, int i) { return i ? v.v.y : v.v.x; }
TL_HD float lane_get(const f4 &v, int i) { return i < 2 ? lane_get(v.a, i) : lane_get(v.b, i - 2); }
TL_HD void lane_set(float &v, int, float s) { v = s; }
TL_HD void lane_set(f2 &v, int i, float s) { if (i) v.v.y = s; else v.v.x = s; }
TL_HD void lane_set(f4 &v, int i, float s) { if (i < 2) lane_set(v.a, i, s); else lane_set(v.b, i - 2, s); }
TL_HD float lane_sum(const float &v) { return v; }
TL_HD float lane_sum(const f2 &v) { return v.v.x + v.v.y; }
TL_HD float lane_sum(const f4 &v) { return (v.a.v.x + v.a.v.y) + (v.b.v.x + v.b.v.y); }
// sum over lanes of p * q, added to acc
TL_HD float lane_dot(const float &p, const float &q, float acc) { return ffma(p, q, acc); }
TL_HD float lane_dot(const f2 &p, const f2 &q, float acc) {
  return ffma(p.v.x, q.v.x, ffma(p.v.y, q.v.y, acc));
}
TL_HD float lane_dot(const f4 &p, const f4 &q, float acc) {
  return lane_dot(p.a, q.a, lane_dot(p.b, q.b, acc));
}

TL_HD double ffma(double a, double b, double c) { return fma(a, b, c); }
TL_HD double frcp(double a) { return 1.0 / a; }
TL_HD double frsqrt(double a) { return 1.0 / sqrt(a); }
TL_HD float fmin2(float a, float b) { return fminf(a, b); }
TL_HD double fmin2(double a, double b) { return fmin(a, b); }

template <class T>
struct Ray {
  T x, y, z, cx, cy, cz;
};

// One surface as a ray of a given wavelength sees it.
struct Surface {
  float c;    // curvature
  float t;    // distance to the next vertex
  float mu;   // n / n'
};

// ---------------------------------------------------------------------------
// EXACT policy.  Statement by statement the reference's eager evaluation.
// ---------------------------------------------------------------------------
TL_HD float exact_cz0(float cx, float cy) {            // rtl:609
  return xsqrt(xsub(xsub(1.0f, xmul(cx, cx)), xmul(cy, cy)));
}

TL_HD void exact_park(bool ok, Ray<float> &r) {         // rtl:574-591
  if (!ok) {
    r.x = 0.f; r.y = 0.f; r.z = 0.f; r.cx = 0.f; r.cy = 0.f; r.cz = 1.f;
  }
}

// Trace one surface.  `count_travel`: this surface takes part in the backward-ray
// test (k > 0 and mask[k-1], rtl:626-628).
TL_HD void exact_surface(Ray<float> &r, const Surface s, bool count_travel, bool allow_backward,
                         bool &ok, bool &backward) {
  // rtl:531-535
  const float e = -xadd(xadd(xmul(r.x, r.cx), xmul(r.y, r.cy)), xmul(r.z, r.cz));
  const float mz = xadd(r.z, xmul(e, r.cz));
  const float m2 = xsub(xadd(xadd(xmul(r.x, r.x), xmul(r.y, r.y)), xmul(r.z, r.z)), xmul(e, e));
  const float temp = xsub(xmul(s.c, m2), xmul(2.0f, mz));
  const float cos2_in = xsub(xmul(r.cz, r.cz), xmul(s.c, temp));
  const bool missed = xsub(cos2_in, kGuard) < 0.0f;                       // rtl:540
  const float cos_in = xsqrt(missed ? 1.0f : cos2_in);                    // rtl:541
  const float dist = xadd(e, xdiv(temp, xadd(r.cz, cos_in)));             // rtl:543
  // rtl:518-521
  const float travel = xmul(dist, r.cz);
  r.x = xadd(r.x, xmul(dist, r.cx));
  r.y = xadd(r.y, xmul(dist, r.cy));
  r.z = xadd(r.z, travel);
  ok = ok && !missed;                                                     // rtl:619
  exact_park(ok, r);
  // rtl:553-568
  const float cos2_out = xsub(1.0f, xmul(xmul(s.mu, s.mu), xsub(1.0f, xmul(cos_in, cos_in))));
  bool lost = xsub(cos2_out, kGuard) < 0.0f;
  const float cos_out = xsqrt(lost ? 1.0f : cos2_out);
  const float g = xsub(cos_out, xmul(s.mu, cos_in));
  const float gc = xmul(g, s.c);
  r.cx = xsub(xmul(s.mu, r.cx), xmul(gc, r.x));
  r.cy = xsub(xmul(s.mu, r.cy), xmul(gc, r.y));
  const float cz2 = xsub(1.0f, xadd(xmul(r.cx, r.cx), xmul(r.cy, r.cy)));
  lost = lost || (xsub(cz2, kGuard) < 0.0f);
  r.cz = xsqrt(lost ? 1.0f : cz2);
  if (count_travel) {                                                     // rtl:626-632
    const bool flagged = (travel < 0.0f) && ok;
    if (allow_backward) backward = backward || flagged;
    else ok = ok && !flagged;
  }
  ok = ok && !lost;                                                       // rtl:635
  exact_park(ok, r);
  r.z = xsub(r.z, s.t);                                                   // rtl:639
}

// Image plane, rtl:660-670.  Returns through r.x, r.y; direction unchanged.
TL_HD void exact_image(Ray<float> &r, bool last_live, bool allow_backward, bool &ok,
                       bool &backward) {
  const float travel = -r.z;
  const float dist = xdiv(travel, r.cz);
  r.x = xadd(r.x, xmul(dist, r.cx));
  r.y = xadd(r.y, xmul(dist, r.cy));
  const bool flagged = (travel < 0.0f) && ok && last_live;
  if (allow_backward) backward = backward || flagged;
  else ok = ok && !flagged;
}

// ---------------------------------------------------------------------------
// FAST policy (generic over the lane type so the CPU check can run it in fp64).
// No masks, no parking: the caller tracks `clear` = the smallest margin by which
// any predicate quantity cleared its threshold and discards the result (falls
// back to the exact policy) unless the ray was clearly good everywhere.
// ---------------------------------------------------------------------------
template <class T>
TL_HD T fast_cz0(T cx, T cy) {
  const T w = ffma(-cy, cy, ffma(-cx, cx, T(1)));
  return w * frsqrt(w);
}

template <class T>
TL_HD void fast_surface(Ray<T> &r, T c, T mu, T mu2, T t, T &min_cos2, T &travel) {
  const T ne = ffma(r.z, r.cz, ffma(r.y, r.cy, r.x * r.cx));          // -e
  const T mz = ffma(-ne, r.cz, r.z);
  const T m2 = ffma(-ne, ne, ffma(r.z, r.z, ffma(r.y, r.y, r.x * r.x)));
  const T tmp = ffma(c, m2, T(-2) * mz);
  const T q = ffma(-c, tmp, r.cz * r.cz);                               // cos^2 in
  const T ci = q * frsqrt(q);
  const T dist = ffma(tmp, frcp(r.cz + ci), -ne);
  travel = dist * r.cz;
  r.x = ffma(dist, r.cx, r.x);
  r.y = ffma(dist, r.cy, r.y);
  r.z = r.z + travel;
  const T qo = ffma(mu2, q, T(1) - mu2);                                // cos^2 out = 1 - mu^2 (1 - q)
  const T co = qo * frsqrt(qo);
  const T gc = ffma(-mu, ci, co) * c;
  r.cx = ffma(-gc, r.x, mu * r.cx);
  r.cy = ffma(-gc, r.y, mu * r.cy);
  const T w = ffma(-r.cy, r.cy, ffma(-r.cx, r.cx, T(1)));
  r.cz = w * frsqrt(w);
  min_cos2 = fmin2(min_cos2, fmin2(q, fmin2(qo, w)));
  r.z = r.z - t;
}

// Image plane; returns the z travel (for the backward-ray margin).
template <class T>
TL_HD T fast_image(Ray<T> &r) {
  const T dist = -r.z * frcp(r.cz);
  r.x = ffma(dist, r.cx, r.x);
  r.y = ffma(dist, r.cy, r.y);
  return -r.z;
}

template <class T>
struct SurfaceGrad {
  T c, t, mu;
};

// ---------------------------------------------------------------------------
// Geometric adjoint (the one the kernels run).
//
// The reference differentiates its scalar formulas; on the manifold of unit
// directions those formulas ARE "intersect the sphere, refract with the unit
// normal" -- so the total derivatives w.r.t. every input equal those of the
// geometric map, which is much cheaper to reverse (no quadratic-solve
// intermediates): with hit point h on F(h) = c|h|^2/2 - h_z = 0, unit normal
// n = z^ - c h, a = n.d, a' = n.d', d' = mu d + g n, g = a' - mu a:
//   transfer  h = r + D d, F(h) = 0   ->  s = -(gh.d)/a, gr = gh + s n, gd += D gr,
//                                          gc -= s |h|^2 / 2        (implicit function theorem)
//   refract                            ->  gg = gd'.n, u = gg/a', ga = -mu g u,
//                                          gmu = gd'.d - u (mu (1 - a^2) + a a'),
//                                          gn = g gd' + ga d, gd = mu gd' + ga n
//   normal    n = z^ - c h             ->  gh -= c gn, gc -= gn.h
//   shift     r' = h - t z^            ->  gt = -gr'_z
// Parked per surface: the hit point (hx, hy) and the incoming direction (dx, dy);
// h_z = c rho / (1 + sqrt(1 - c^2 rho)) (the hit lies on the sphere; 1 - c h_z = sqrt(...) is
// also n_z) and d_z = sqrt(1 - dx^2 - dy^2) are rebuilt.  Verified against autograd of the
// oracle in fp64 (tests/test_core_cpu.py, ~1e-12).
// ---------------------------------------------------------------------------
template <class T>
struct Vec3 {
  T x, y, z;
};

template <class T>
TL_HD T dot3(const Vec3<T> &p, const Vec3<T> &q) {
  return ffma(p.z, q.z, ffma(p.y, q.y, p.x * q.x));
}

// State carried from surface k+1 to surface k of the sweep.
template <class T>
struct Sweep {
  Vec3<T> hit;   // hit point on surface k+1 (its vertex coordinates; image plane: z = 0)
  Vec3<T> dir;   // direction of the ray between surfaces k and k+1
  Vec3<T> gr;    // adjoint of a point of that ray
  Vec3<T> gd;    // adjoint of its direction, still missing the (distance to `hit`) * gr term
};

// Image plane: seeds on the image point (gx, gy) and on the final direction (gcx, gcy).
// `pre` = state in front of the image plane, (x_img, y_img) = the traced image point.
template <class T>
TL_HD Sweep<T> sweep_begin(const Ray<T> &pre, T x_img, T y_img, T gx, T gy, T gcx, T gcy) {
  Sweep<T> s;
  s.hit = Vec3<T>{x_img, y_img, T(0)};
  s.dir = Vec3<T>{pre.cx, pre.cy, pre.cz};
  s.gr = Vec3<T>{gx, gy, -ffma(gy, pre.cy, gx * pre.cx) * frcp(pre.cz)};   // plane: n = z^
  s.gd = Vec3<T>{gcx, gcy, T(0)};
  return s;
}

template <class T>
TL_HD SurfaceGrad<T> sweep_sphere(Sweep<T> &s, T hx, T hy, T dx, T dy, T c, T t, T mu, T mu2) {
  SurfaceGrad<T> g;
  // rebuild the dependent components
  const T rho = ffma(hy, hy, hx * hx);
  const T w = ffma(-(c * c), rho, T(1));
  const T root = w * frsqrt(w);                       // = 1 - c h_z = n_z
  const T hz = (c * rho) * frcp(T(1) + root);
  const T wd = ffma(-dy, dy, ffma(-dx, dx, T(1)));
  const T dz = wd * frsqrt(wd);
  // finish the transfer behind this surface: distance from this hit to the next one
  const T dist = ffma((s.hit.z - hz) + t, s.dir.z, ffma(s.hit.y - hy, s.dir.y, (s.hit.x - hx) * s.dir.x));
  const Vec3<T> gdo{ffma(dist, s.gr.x, s.gd.x), ffma(dist, s.gr.y, s.gd.y), ffma(dist, s.gr.z, s.gd.z)};
  g.t = -s.gr.z;
  // refraction d' = mu d + g n
  const Vec3<T> n{-c * hx, -c * hy, root};
  const Vec3<T> d{dx, dy, dz};
  const T a = dot3(n, d);
  const T ap = dot3(n, s.dir);
  const T gsn = ffma(-mu, a, ap);
  const T gdd = dot3(gdo, d);
  const T u = dot3(gdo, n) * frcp(ap);
  const T ga = -(mu * gsn) * u;
  g.mu = ffma(-u, ffma(a, ap, mu * ffma(-a, a, T(1))), gdd);
  const Vec3<T> gn{ffma(ga, d.x, gsn * gdo.x), ffma(ga, d.y, gsn * gdo.y), ffma(ga, d.z, gsn * gdo.z)};
  const Vec3<T> gdi{ffma(ga, n.x, mu * gdo.x), ffma(ga, n.y, mu * gdo.y), ffma(ga, n.z, mu * gdo.z)};
  // normal n = z^ - c h
  const Vec3<T> gh{ffma(-c, gn.x, s.gr.x), ffma(-c, gn.y, s.gr.y), ffma(-c, gn.z, s.gr.z)};
  const T gc_n = ffma(gn.z, hz, ffma(gn.y, hy, gn.x * hx));
  // transfer onto the sphere
  const T sd = -dot3(gh, d) * frcp(a);
  s.gr = Vec3<T>{ffma(sd, n.x, gh.x), ffma(sd, n.y, gh.y), ffma(sd, n.z, gh.z)};
  g.c = -ffma(sd * T(0.5), ffma(hz, hz, rho), gc_n);
  s.gd = gdi;
  s.hit = Vec3<T>{hx, hy, hz};
  s.dir = d;
  return g;
}

// Entrance: the ray starts at (x, y, z_in) with direction (cx, cy, sqrt(1 - cx^2 - cy^2)).
template <class T>
TL_HD void sweep_end(const Sweep<T> &s, T z_in, T &gx, T &gy, T &gz, T &gcx, T &gcy) {
  const T rdz = frcp(s.dir.z);
  const T dist = (s.hit.z - z_in) * rdz;
  const T gdz = ffma(dist, s.gr.z, s.gd.z) * rdz;
  gx = s.gr.x;
  gy = s.gr.y;
  gz = s.gr.z;
  gcx = ffma(-gdz, s.dir.x, ffma(dist, s.gr.x, s.gd.x));     // cz is a function of (cx, cy)
  gcy = ffma(-gdz, s.dir.y, ffma(dist, s.gr.y, s.gd.y));
}

}  // namespace tl
