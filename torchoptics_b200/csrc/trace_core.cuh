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

// Forward-mode pair: a value with its derivatives along two input directions (the pupil's x and
// y).  fast_surface / fast_image instantiate with it unchanged, which is how the on-device ray
// aiming (rtl:129-208) gets d(stop point)/d(pupil point) of its three 'tee' rays without a
// backward pass.
struct D2 {
  float v, a, b;
  TL_HD D2() {}
  TL_HD D2(float s) : v(s), a(0.f), b(0.f) {}
  TL_HD D2(float v_, float a_, float b_) : v(v_), a(a_), b(b_) {}
};
TL_HD D2 operator-(D2 p) { return D2(-p.v, -p.a, -p.b); }
TL_HD D2 operator+(D2 p, D2 q) { return D2(p.v + q.v, p.a + q.a, p.b + q.b); }
TL_HD D2 operator-(D2 p, D2 q) { return D2(p.v - q.v, p.a - q.a, p.b - q.b); }
TL_HD D2 operator*(D2 p, D2 q) {
  return D2(p.v * q.v, ffma(p.a, q.v, p.v * q.a), ffma(p.b, q.v, p.v * q.b));
}
TL_HD D2 ffma(D2 p, D2 q, D2 r) {
  return D2(ffma(p.v, q.v, r.v), ffma(p.a, q.v, ffma(p.v, q.a, r.a)), ffma(p.b, q.v, ffma(p.v, q.b, r.b)));
}
TL_HD D2 frcp(D2 p) {
  const float r = 1.0f / p.v, d = -r * r;
  return D2(r, d * p.a, d * p.b);
}
TL_HD D2 frsqrt(D2 p) {
  const float r = 1.0f / sqrtf(p.v), d = -0.5f * r / p.v;
  return D2(r, d * p.a, d * p.b);
}
TL_HD D2 fmin2(D2 p, D2 q) { return p.v <= q.v ? p : q; }

template <class T>
struct Ray {
  T x, y, z, cx, cy, cz;
};

// What the reversible adjoint sweep (sweep_sphere_rev below) needs from the forward pass, per
// ray-surface event.
template <class T>
struct Parked {
  T dist;   // marching distance from the previous vertex-shifted point to this surface (rtl:543)
  T ci;     // cos(theta)  at this surface (rtl:541)
  T co;     // cos(theta') at this surface (rtl:556)
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

// Penalty terms of trace_skew(aggregate=True) for one ray at one surface, rtl:641-657:
// z_RELU, theta_norm, theta_prime_norm.
struct Penalty {
  float z_relu, theta, theta_prime;
};

constexpr float kClampCos = 1.0f - 1e-7f;          // rtl:645-647: clamp(.., max=1-1e-7) in fp32
constexpr float kHalfPi = 1.5707963267948966f;     // rtl:651: theta / (1/2*pi)

// acos(clamp(sqrt(cos2), -1+1e-7, 1-1e-7)) / (pi/2), rtl:646-652 (a NaN from a negative cos2 only
// occurs on failed rays, whose angle is overwritten with 1, rtl:653-654)
TL_HD float exact_angle_norm(float cos2) {
  float v = xsqrt(cos2);
  v = v > kClampCos ? kClampCos : v;
  v = v < -kClampCos ? -kClampCos : v;
  return xdiv(acosf(v), kHalfPi);
}

// Trace one surface.  `count_travel`: this surface takes part in the backward-ray
// test (k > 0 and mask[k-1], rtl:626-628).  PEN: also return the aggregate=True terms.
template <bool PEN>
TL_HD void exact_surface_t(Ray<float> &r, const Surface s, bool count_travel, bool allow_backward,
                           bool &ok, bool &backward, Penalty *pen, Parked<float> *pk = nullptr) {
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
  if (pk) {
    pk->dist = dist;
    pk->ci = cos_in;
    pk->co = cos_out;
  }
  if (PEN) {                                                              // rtl:641-657
    pen->z_relu = r.z <= 0.0f ? 0.0f : r.z;
    pen->theta = ok ? exact_angle_norm(cos2_in) : 1.0f;
    pen->theta_prime = ok ? exact_angle_norm(cos2_out) : 1.0f;
  }
}

TL_HD void exact_surface(Ray<float> &r, const Surface s, bool count_travel, bool allow_backward,
                         bool &ok, bool &backward) {
  exact_surface_t<false>(r, s, count_travel, allow_backward, ok, backward, nullptr);
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

// One surface, fast policy.  Round 2: the marching distance is the near root of
//     c s^2 - 2 beta s + gamma = 0,   beta = cz - c (r.d),   gamma = c |r|^2 - 2 z,
// whose discriminant beta^2 - c gamma IS cos^2(theta) of rtl:535 (expand both: identical
// polynomials) -- one operation shorter than the e / mz / m2 / temp chain of rtl:531-543 and without
// its cancellation between `e` and `temp / (cz + cos)` for a ray that starts close to the surface.
// Every fast-path kernel shares this arithmetic, so a loss evaluated with and without gradients, fused
// or split, sees bit-identical ray heights.  (`Parked`: what the reversible adjoint needs, below.)
template <class T>
TL_HD void fast_surface_core(Ray<T> &r, T c, T mu, T mu2, T om2, T t, T &min_cos2, T &travel, Parked<T> &pk) {
  const T rd = ffma(r.z, r.cz, ffma(r.y, r.cy, r.x * r.cx));            // r.d
  const T beta = ffma(-c, rd, r.cz);
  const T r2 = ffma(r.z, r.z, ffma(r.y, r.y, r.x * r.x));
  const T gamma = ffma(c, r2, T(-2) * r.z);
  const T q = ffma(-c, gamma, beta * beta);                              // cos^2 in
  const T ci = q * frsqrt(q);
  const T dist = gamma * frcp(beta + ci);
  travel = dist * r.cz;
  r.x = ffma(dist, r.cx, r.x);
  r.y = ffma(dist, r.cy, r.y);
  r.z = r.z + travel;
  const T qo = ffma(mu2, q, om2);                                        // cos^2 out, om2 = 1 - mu^2
  const T co = qo * frsqrt(qo);
  const T g = ffma(-mu, ci, co);
  const T gc = g * c;
  r.cx = ffma(-gc, r.x, mu * r.cx);
  r.cy = ffma(-gc, r.y, mu * r.cy);
  const T w = ffma(-r.cy, r.cy, ffma(-r.cx, r.cx, T(1)));                // rtl:566, renormalised
  r.cz = w * frsqrt(w);
  min_cos2 = fmin2(min_cos2, fmin2(q, fmin2(qo, w)));
  r.z = r.z - t;
  pk.dist = dist;
  pk.ci = ci;
  pk.co = co;
}

template <class T>
TL_HD void fast_surface(Ray<T> &r, T c, T mu, T mu2, T t, T &min_cos2, T &travel, T &cos_in,
                        T &cos_out) {
  Parked<T> pk;
  fast_surface_core(r, c, mu, mu2, T(1) - mu2, t, min_cos2, travel, pk);
  cos_in = pk.ci;
  cos_out = pk.co;
}

template <class T>
TL_HD void fast_surface(Ray<T> &r, T c, T mu, T mu2, T t, T &min_cos2, T &travel) {
  T cos_in, cos_out;
  fast_surface(r, c, mu, mu2, t, min_cos2, travel, cos_in, cos_out);
}

// Fast-policy penalty terms of a ray that is clear of every threshold (so it is ok): the angles
// from the cosines the trace already has, z_RELU from the shifted z.
// acos on [-1, 1] as sqrt(1 - |x|) * P7(|x|) (Abramowitz & Stegun 4.4.46, |error| <= 2e-8 rad):
// 7 FMAs and one square root instead of the ~25 instructions of acosf -- the fast-policy angle
// terms are issue-bound on it (two per ray-surface event).  1 - |x| is exact near |x| = 1
// (Sterbenz), so small angles keep their relative accuracy to ~1e-7 + 2e-8 / theta.
TL_HD float fast_acos(float x) {
  const float ax = fabsf(x);
  float p = -0.0012624911f;
  p = ffma(p, ax, 0.0066700901f);
  p = ffma(p, ax, -0.0170881256f);
  p = ffma(p, ax, 0.0308918810f);
  p = ffma(p, ax, -0.0501743046f);
  p = ffma(p, ax, 0.0889789874f);
  p = ffma(p, ax, -0.2145988016f);
  p = ffma(p, ax, 1.5707963050f);
  const float t = fmaxf(1.0f - ax, 0.0f);
#if defined(__CUDA_ARCH__)
  float root;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(root) : "f"(t));
#else
  const float root = sqrtf(t);
#endif
  const float r = root * p;
  return x < 0.0f ? 3.14159265358979f - r : r;
}
TL_HD float fast_angle_norm(float cosv) {
  return fast_acos(fminf(cosv, kClampCos)) * (1.0f / kHalfPi);
}
TL_HD double fast_angle_norm(double cosv) {      // fp64 check build: the clamp bound in fp64
  return acos(fmin(cosv, 1.0 - 1e-7)) * (1.0 / 1.5707963267948966);
}
// sin^2 of the clamp angle: cos > bound  <=>  sin^2 < 1 - bound^2 (fp32 bound 1 - 2^-23; the fp64
// check build uses the fp64 bound 1 - 1e-7)
template <class T> TL_HD T clamp_sin2() { return T(2.3841856e-07f); }
template <> TL_HD double clamp_sin2<double>() { return 1.0 - (1.0 - 1e-7) * (1.0 - 1e-7); }

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

// ---------------------------------------------------------------------------
// REVERSIBLE formulation (round 2; what the fused spot kernel k_spot_rev runs).
//
// A traced ray can be walked backwards: refraction and transfer are invertible maps.  The adjoint
// sweep therefore does not need the hit points and directions of the forward pass -- it rebuilds
// them surface by surface on its way back from the image plane,
//     h_k = h_{k+1} + t_k z^ - D_{k+1} d'_k             (undo the transfer)
//     d_k = (d'_k - g_k n_k) / mu_k,  g_k = a'_k - mu_k a_k   (undo the refraction)
// from the three scalars per ray-surface event that are NOT cheap to recompute: the marching
// distance D_k and the two cosines a_k = cos(theta), a'_k = cos(theta') (each is behind a square
// root).  Against parking (hx, hy, dx, dy) and rebuilding h_z, d_z, n.d, n.d' this is 12 B instead
// of 16 B per event of parked state, 2 instead of 5 MUFU and ~61 instead of ~72 FMA-pipe
// operations per event in the sweep -- and n_z = 1 - c h_z comes from the true h_z, so a hit
// beyond the equator of the sphere needs no branch bit.
//
// The forward is fast_surface_core (every fast-path kernel's), handing out what it computed anyway.
// (Tried and dropped: cz' = mu cz + g n_z instead of the renormalising sqrt of rtl:566 saves a MUFU,
// but |d| then drifts by an ULP per surface and the RMS of an 8x8 pupil moved by up to 1.2e-5 of
// itself against 5e-6 for the renormalised form -- outside the 1e-5 budget.)
// ---------------------------------------------------------------------------
template <class T>
TL_HD void fast_surface_rev(Ray<T> &r, T c, T mu, T mu2, T om2, T t, T &min_cos2, T &min_cz,
                            T &travel, Parked<T> &pk) {
  (void)min_cz;
  fast_surface_core(r, c, mu, mu2, om2, t, min_cos2, travel, pk);
}

// cz' must clear this to stay on the fast path: sqrt(kGuard + kBandCos2), rounded up
constexpr float kBandCz = 0.01006f;

template <class T>
struct SweepRev {
  Vec3<T> hit;   // hit point on surface k+1 (its vertex coordinates; image plane: z = 0)
  Vec3<T> dir;   // direction of the ray between surfaces k and k+1
  Vec3<T> gr;    // adjoint of a point of that ray
  Vec3<T> gd;    // adjoint of its direction, still missing the (distance to `hit`) * gr term
  T dnext;       // distance from the hit on surface k to `hit`
};

template <class T>
TL_HD SweepRev<T> sweep_begin_rev(const Ray<T> &pre, T x_img, T y_img, T gx, T gy, T gcx, T gcy) {
  SweepRev<T> s;
  const T rcz = frcp(pre.cz);
  s.hit = Vec3<T>{x_img, y_img, T(0)};
  s.dir = Vec3<T>{pre.cx, pre.cy, pre.cz};
  s.gr = Vec3<T>{gx, gy, -ffma(gy, pre.cy, gx * pre.cx) * rcz};         // plane: n = z^
  s.gd = Vec3<T>{gcx, gcy, T(0)};
  s.dnext = -pre.z * rcz;
  return s;
}

// One surface of the backward walk.  a = cos(theta), ap = cos(theta'), rap = 1 / ap, rmu = 1 / mu;
// `dist` = the parked marching distance of THIS surface (handed on to the next step).
template <class T>
TL_HD SurfaceGrad<T> sweep_sphere_rev_core(SweepRev<T> &s, T dist, T a, T ap, T rap, T c, T t, T mu, T rmu) {
  SurfaceGrad<T> g;
  const T dn = s.dnext;
  // undo the transfer: the hit point on this surface
  const Vec3<T> h{ffma(-dn, s.dir.x, s.hit.x), ffma(-dn, s.dir.y, s.hit.y), ffma(-dn, s.dir.z, s.hit.z + t)};
  const Vec3<T> gdo{ffma(dn, s.gr.x, s.gd.x), ffma(dn, s.gr.y, s.gd.y), ffma(dn, s.gr.z, s.gd.z)};
  g.t = -s.gr.z;
  const Vec3<T> n{-c * h.x, -c * h.y, ffma(-c, h.z, T(1))};
  const T gsn = ffma(-mu, a, ap);
  // undo the refraction: the incoming direction d = (d' - g n) / mu
  const T grm = gsn * rmu;
  const Vec3<T> d{ffma(-grm, n.x, rmu * s.dir.x), ffma(-grm, n.y, rmu * s.dir.y), ffma(-grm, n.z, rmu * s.dir.z)};
  const T gdd = dot3(gdo, d);
  const T u = dot3(gdo, n) * rap;
  const T ga = -(mu * gsn) * u;
  g.mu = ffma(-u, ffma(a, ap, mu * ffma(-a, a, T(1))), gdd);
  const Vec3<T> gn{ffma(ga, d.x, gsn * gdo.x), ffma(ga, d.y, gsn * gdo.y), ffma(ga, d.z, gsn * gdo.z)};
  const Vec3<T> gdi{ffma(ga, n.x, mu * gdo.x), ffma(ga, n.y, mu * gdo.y), ffma(ga, n.z, mu * gdo.z)};
  const Vec3<T> gh{ffma(-c, gn.x, s.gr.x), ffma(-c, gn.y, s.gr.y), ffma(-c, gn.z, s.gr.z)};
  const T gc_n = dot3(gn, h);
  const T sd = -dot3(gh, d) * frcp(a);
  s.gr = Vec3<T>{ffma(sd, n.x, gh.x), ffma(sd, n.y, gh.y), ffma(sd, n.z, gh.z)};
  g.c = -ffma(sd * T(0.5), dot3(h, h), gc_n);
  s.gd = gdi;
  s.hit = h;
  s.dir = d;
  s.dnext = dist;
  return g;
}

// ... from all three parked values
template <class T>
TL_HD SurfaceGrad<T> sweep_sphere_rev(SweepRev<T> &s, const Parked<T> &pk, T c, T t, T mu, T rmu) {
  return sweep_sphere_rev_core(s, pk.dist, pk.ci, pk.co, frcp(pk.co), c, t, mu, rmu);
}

// ... from (dist, cos theta) alone: cos theta' = sqrt(1 - mu^2 (1 - cos^2 theta)) (rtl:553-556) is
// rebuilt -- three more FMA-pipe operations per event, but its rsqrt IS the 1 / cos theta' the sweep
// needs anyway (no extra MUFU) and the parked state shrinks to 8 bytes per event.
template <class T>
TL_HD SurfaceGrad<T> sweep_sphere_rev2(SweepRev<T> &s, T dist, T ci, T c, T t, T mu, T mu2, T om2, T rmu) {
  const T qo = ffma(mu2, ci * ci, om2);
  const T rap = frsqrt(qo);
  return sweep_sphere_rev_core(s, dist, ci, qo * rap, rap, c, t, mu, rmu);
}

// Entrance: the ray starts at (x, y, z_in) with direction (cx, cy, sqrt(1 - cx^2 - cy^2)); the
// distance to the first hit is the parked one.
template <class T>
TL_HD void sweep_end_rev(const SweepRev<T> &s, T &gx, T &gy, T &gz, T &gcx, T &gcy) {
  const T rdz = frcp(s.dir.z);
  const T dist = s.dnext;
  const T gdz = ffma(dist, s.gr.z, s.gd.z) * rdz;
  gx = s.gr.x;
  gy = s.gr.y;
  gz = s.gr.z;
  gcx = ffma(-gdz, s.dir.x, ffma(dist, s.gr.x, s.gd.x));     // cz is a function of (cx, cy)
  gcy = ffma(-gdz, s.dir.y, ffma(dist, s.gr.y, s.gd.y));
}

// Did the ray hit the sphere beyond its equator?  `z_shifted` = h_z - t (the state behind the
// surface, rtl:639).
TL_HD bool beyond_equator(float c, float z_shifted, float t) { return c * (z_shifted + t) > 1.0f; }

// Lane-wise selects for the penalty seeds.
TL_HD float keep_if(bool cond, float v) { return cond ? v : 0.0f; }
TL_HD double keep_if(bool cond, double v) { return cond ? v : 0.0; }
TL_HD float keep_pos(float cond, float v) { return cond > 0.0f ? v : 0.0f; }
TL_HD double keep_pos(double cond, double v) { return cond > 0.0 ? v : 0.0; }
TL_HD f2 keep_pos(f2 cond, f2 v) { return f2(keep_pos(cond.v.x, v.v.x), keep_pos(cond.v.y, v.v.y)); }
TL_HD f4 keep_pos(f4 cond, f4 v) { return f4(keep_pos(cond.a, v.a), keep_pos(cond.b, v.b)); }

// sweep_sphere with the seeds of the aggregate=True terms of this surface (rtl:641-657):
//   pz   on z_RELU = relu(h_z - t)            -> joins the adjoint of the point behind the surface
//   pth  on theta_norm  = acos(min(a, 1-1e-7)) / (pi/2),   a  = n.d  = cos(theta)
//   pthp on theta_prime_norm, same of a' = sqrt(1 - mu^2 (1 - a^2)) = cos(theta')
// The two angle seeds enter where the refraction's own adjoint of a enters (ga) plus a direct
// term on mu.  Seeds must already be zero for lanes that are not ok behind this surface.
// `branch` = +1, or -1 for a hit BEYOND THE EQUATOR of the sphere (c h_z > 1, possible for very
// oblique rays): there n_z = 1 - c h_z = -sqrt(1 - c^2 rho), which the parked (hx, hy) alone cannot
// tell.  (sweep_sphere assumes +1: a ray with such a hit that still reaches the image is outside
// its validity; the penalty terms exist to punish exactly such rays, so this one carries the sign.)
// Also hands back the three quantities the terms are made of -- cos(theta), cos(theta') as the
// reference forms them and z' = h_z - t -- for callers that accumulate the penalty value itself.
template <class T>
TL_HD SurfaceGrad<T> sweep_sphere_pen(Sweep<T> &s, T hx, T hy, T dx, T dy, T c, T t, T mu, T mu2,
                                      T pz, T pth, T pthp, T branch, T &cos_in, T &cos_out,
                                      T &z_behind) {
  SurfaceGrad<T> g;
  const T rho = ffma(hy, hy, hx * hx);
  const T w = ffma(-(c * c), rho, T(1));
  const T root = branch * (w * frsqrt(w));
  const T hz = (c * rho) * frcp(T(1) + root);
  const T wd = ffma(-dy, dy, ffma(-dx, dx, T(1)));
  const T dz = wd * frsqrt(wd);
  const T dist = ffma((s.hit.z - hz) + t, s.dir.z, ffma(s.hit.y - hy, s.dir.y, (s.hit.x - hx) * s.dir.x));
  const Vec3<T> gdo{ffma(dist, s.gr.x, s.gd.x), ffma(dist, s.gr.y, s.gd.y), ffma(dist, s.gr.z, s.gd.z)};
  // z_RELU = relu(z'), z' = h_z - t: a seed on the START point of the ray behind the surface (it
  // joins the point adjoint after the direction adjoint above has taken its dist * gr share)
  s.gr.z = s.gr.z + keep_pos(hz - t, pz);
  g.t = -s.gr.z;
  const Vec3<T> n{-c * hx, -c * hy, root};
  const Vec3<T> d{dx, dy, dz};
  const T a = dot3(n, d);
  const T ap = dot3(n, s.dir);
  const T gsn = ffma(-mu, a, ap);
  const T gdd = dot3(gdo, d);
  const T rap = frcp(ap);
  const T u = dot3(gdo, n) * rap;
  // angle seeds: d theta / d a = -(2/pi) / sqrt(1 - a^2) inside the clamp, 0 outside
  // cos(theta') as the reference forms it, sqrt(1 - mu^2 (1 - a^2)) (rtl:553): equal to n.d' for a
  // physical ray, but not when the reference's always-positive cz' (rtl:566) has flipped d'
  // sin^2(theta) = |n x d|^2 rather than 1 - a^2: near normal incidence, where d theta / d a =
  // 1 / sin(theta) amplifies everything, 1 - a^2 has lost its digits to cancellation
  const T kx = ffma(n.y, d.z, -(n.z * d.y)), ky = ffma(n.z, d.x, -(n.x * d.z)), kz = ffma(n.x, d.y, -(n.y * d.x));
  const T sin2 = ffma(kz, kz, ffma(ky, ky, kx * kx));
  const T sin2p = mu2 * sin2;
  const T cos2p = T(1) - sin2p;
  const T rapp = frsqrt(cos2p);                       // 1 / cos(theta')
  cos_in = a;
  cos_out = cos2p * rapp;
  z_behind = hz - t;
  const T clamp_s2 = clamp_sin2<T>();
  // (torch's clamp passes the gradient on the boundary: keep cos <= bound, drop cos > bound -- decided
  // on the well-conditioned sin^2, not on a cosine that is within an ULP of 1 -- and keep the rsqrt
  // argument away from 0 on the dropped lanes)
  const T over = keep_pos(clamp_s2 - sin2, T(1)), overp = keep_pos(clamp_s2 - sin2p, T(1));
  const T dth = (pth - over * pth) * frsqrt(sin2 + over);
  const T dthp = (pthp - overp * pthp) * frsqrt(sin2p + overp);
  const T gap = T(-0.63661977236758134) * dthp;                           // adjoint of a'
  const T ga = ffma(gap * mu2, a * rapp, ffma(T(-0.63661977236758134), dth, -(mu * gsn) * u));
  g.mu = ffma(-(gap * mu), sin2 * rapp, ffma(-u, ffma(a, ap, mu * sin2), gdd));
  const Vec3<T> gn{ffma(ga, d.x, gsn * gdo.x), ffma(ga, d.y, gsn * gdo.y), ffma(ga, d.z, gsn * gdo.z)};
  const Vec3<T> gdi{ffma(ga, n.x, mu * gdo.x), ffma(ga, n.y, mu * gdo.y), ffma(ga, n.z, mu * gdo.z)};
  const Vec3<T> gh{ffma(-c, gn.x, s.gr.x), ffma(-c, gn.y, s.gr.y), ffma(-c, gn.z, s.gr.z)};
  const T gc_n = ffma(gn.z, hz, ffma(gn.y, hy, gn.x * hx));
  const T sd = -dot3(gh, d) * frcp(a);
  s.gr = Vec3<T>{ffma(sd, n.x, gh.x), ffma(sd, n.y, gh.y), ffma(sd, n.z, gh.z)};
  g.c = -ffma(sd * T(0.5), ffma(hz, hz, rho), gc_n);
  s.gd = gdi;
  s.hit = Vec3<T>{hx, hy, hz};
  s.dir = d;
  return g;
}

template <class T>
TL_HD SurfaceGrad<T> sweep_sphere_pen(Sweep<T> &s, T hx, T hy, T dx, T dy, T c, T t, T mu, T mu2,
                                      T pz, T pth, T pthp, T branch) {
  T cos_in, cos_out, z_behind;
  return sweep_sphere_pen(s, hx, hy, dx, dy, c, t, mu, mu2, pz, pth, pthp, branch, cos_in, cos_out, z_behind);
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
