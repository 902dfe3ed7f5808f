// Per-ray arithmetic of the EXTENSION surfaces: conic /
// even-asphere sag solved by Newton iteration, vector Snell with the general normal,
// clear semi-diameter clip, optical path length; and the geometric adjoint of such a
// surface.  The reference has none of this ("parity unpinned"): behaviour is defined by
// oracle/asphere_oracle.py, whose statement order the exact policy below follows.
//
//   s(rho) = c rho / (1 + sqrt(1 - (1+k) c^2 rho)) + sum_{i=2..8} a_i rho^i ,  rho = x^2 + y^2
#pragma once
#include "trace_core.cuh"

namespace tl {

constexpr int kAsphCoefs = 7;     // a4, a6, ..., a16
constexpr int kNewton = 4;        // fixed iteration count (oracle: N_NEWTON)
// Work per asphere event at kNewton = 4 (FMA = 2 flop; DESIGN.md section 7b has the table):
//   algorithmic  61 + 55 * 4 + 35 = 316 forward, 105 + 2 * 55 + 35 = 250 adjoint, 566 together
//   executed     fast_asph_surface 26 (start) + 4 * 57 (iterations) + 94 (hit, normal, Snell, margins) = 348,
//                19 MUFU; sweep_asphere 238 + 33 (sums) = 271, 8 MUFU
//                (the fast policy leaves the loop early, newton_settled below: 2 iterations per event on the
//                config-3 lens, 348 -> ~234 forward, less where the evaluation behind the loop is skipped too,
//                newton_is_fresh; the exact policy always runs the four)
constexpr int kAsphParams = 2 + kAsphCoefs;   // c, k, a4..a16

template <class S>
struct AsphSurfaceT {
  S c, k, a[kAsphCoefs], t, mu, sd2;   // sd2 = (clear semi-diameter)^2, +inf = no clip
};
using AsphSurface = AsphSurfaceT<float>;   // (the fp64 instantiation exists for the CPU gradient check)

// ---------------------------------------------------------------------------
// EXACT policy (scalar, individually rounded, oracle statement order)
// ---------------------------------------------------------------------------
TL_HD void exact_poly(const float *a, float rho, float &poly, float &dpoly) {
  float inner = a[6];
  float d_inner = xmul(8.0f, a[6]);
  const float mult[6] = {7.f, 6.f, 5.f, 4.f, 3.f, 2.f};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < 6; ++j) {
    const float coef = a[5 - j];
    inner = xadd(xmul(inner, rho), coef);
    d_inner = xadd(xmul(d_inner, rho), xmul(mult[j], coef));
  }
  poly = xmul(xmul(inner, rho), rho);
  dpoly = xmul(d_inner, rho);
}

TL_HD void exact_sag_and_slope(const AsphSurface &s, float rho, float &sag, float &slope,
                               float &radicand) {
  radicand = xsub(1.0f, xmul(xmul(xmul(xadd(1.0f, s.k), s.c), s.c), rho));
  const float safe = (xsub(radicand, kGuard) < 0.0f) ? 1.0f : radicand;
  const float root = xsqrt(safe);
  float poly, dpoly;
  exact_poly(s.a, rho, poly, dpoly);
  sag = xadd(xdiv(xmul(s.c, rho), xadd(1.0f, root)), poly);
  slope = xadd(xdiv(s.c, xmul(2.0f, root)), dpoly);
}

// One general surface, exact policy.  `index` = refractive index in front of the surface,
// `opl` accumulates the optical path.
TL_HD void exact_asph_surface(Ray<float> &r, const AsphSurface &s, bool count_travel,
                              bool allow_backward, bool &ok, bool &backward, float &index,
                              float &opl) {
  // base-sphere start, rtl:531-543
  const float e = -xadd(xadd(xmul(r.x, r.cx), xmul(r.y, r.cy)), xmul(r.z, r.cz));
  const float mz = xadd(r.z, xmul(e, r.cz));
  const float m2 = xsub(xadd(xadd(xmul(r.x, r.x), xmul(r.y, r.y)), xmul(r.z, r.z)), xmul(e, e));
  const float temp = xsub(xmul(s.c, m2), xmul(2.0f, mz));
  const float cos2_base = xsub(xmul(r.cz, r.cz), xmul(s.c, temp));
  const bool missed = xsub(cos2_base, kGuard) < 0.0f;
  const float cos_base = xsqrt(missed ? 1.0f : cos2_base);
  float tau = xadd(e, xdiv(temp, xadd(r.cz, cos_base)));
  if (missed) tau = 0.0f;
  for (int it = 0; it < kNewton; ++it) {
    const float hx = xadd(r.x, xmul(tau, r.cx));
    const float hy = xadd(r.y, xmul(tau, r.cy));
    const float rho = xadd(xmul(hx, hx), xmul(hy, hy));
    float sag, slope, radicand;
    exact_sag_and_slope(s, rho, sag, slope, radicand);
    const float f = xsub(xadd(r.z, xmul(tau, r.cz)), sag);
    const float fp = xsub(r.cz, xmul(xmul(slope, 2.0f), xadd(xmul(hx, r.cx), xmul(hy, r.cy))));
    tau = xsub(tau, xdiv(f, fp));
  }
  const float travel = xmul(tau, r.cz);
  r.x = xadd(r.x, xmul(tau, r.cx));
  r.y = xadd(r.y, xmul(tau, r.cy));
  r.z = xadd(r.z, travel);
  const float rho = xadd(xmul(r.x, r.x), xmul(r.y, r.y));
  float sag, slope, radicand;
  exact_sag_and_slope(s, rho, sag, slope, radicand);
  const bool finite = fabsf(tau) <= 3.4028234e38f;       // false for NaN and inf
  const bool failed = missed || (xsub(radicand, kGuard) < 0.0f) || (rho > s.sd2) || !finite;
  ok = ok && !failed;
  if (ok) opl = xadd(opl, xmul(index, tau));
  exact_park(ok, r);
  if (!ok) slope = 0.0f;
  // unit normal and vector Snell
  float nx = xmul(xmul(-2.0f, r.x), slope);
  float ny = xmul(xmul(-2.0f, r.y), slope);
  const float inv_norm = xdiv(1.0f, xsqrt(xadd(xadd(xmul(nx, nx), xmul(ny, ny)), 1.0f)));
  nx = xmul(nx, inv_norm);
  ny = xmul(ny, inv_norm);
  const float nz = inv_norm;
  const float cos_in = xadd(xadd(xmul(r.cx, nx), xmul(r.cy, ny)), xmul(r.cz, nz));
  const float cos2_out = xsub(1.0f, xmul(xmul(s.mu, s.mu), xsub(1.0f, xmul(cos_in, cos_in))));
  bool lost = xsub(cos2_out, kGuard) < 0.0f;
  const float cos_out = xsqrt(lost ? 1.0f : cos2_out);
  const float g = xsub(cos_out, xmul(s.mu, cos_in));
  r.cx = xadd(xmul(s.mu, r.cx), xmul(g, nx));
  r.cy = xadd(xmul(s.mu, r.cy), xmul(g, ny));
  const float cz2 = xsub(1.0f, xadd(xmul(r.cx, r.cx), xmul(r.cy, r.cy)));
  lost = lost || (xsub(cz2, kGuard) < 0.0f);
  r.cz = xsqrt(lost ? 1.0f : cz2);
  if (count_travel) {
    const bool flagged = (travel < 0.0f) && ok;
    if (allow_backward) backward = backward || flagged;
    else ok = ok && !flagged;
  }
  ok = ok && !lost;
  exact_park(ok, r);
  r.z = xsub(r.z, s.t);
  index = xdiv(index, s.mu);
}

TL_HD void exact_asph_image(Ray<float> &r, bool last_live, bool allow_backward, bool &ok,
                            bool &backward, float index, float &opl) {
  const float travel = -r.z;
  const float dist = xdiv(travel, r.cz);
  r.x = xadd(r.x, xmul(dist, r.cx));
  r.y = xadd(r.y, xmul(dist, r.cy));
  if (ok) opl = xadd(opl, xmul(index, dist));
  const bool flagged = (travel < 0.0f) && ok && last_live;
  if (allow_backward) backward = backward || flagged;
  else ok = ok && !flagged;
}

// ---------------------------------------------------------------------------
// FAST policy (generic lane type)
// ---------------------------------------------------------------------------
template <class T>
struct AsphEval {
  T sag, slope, curv;   // s, ds/drho, d2s/drho2
  T radicand, rs;       // 1 - (1+k) c^2 rho and its reciprocal square root
  T inv1r;              // 1 / (1 + sqrt(radicand))
};

// s, s', (s'' when CURV) at rho.  Uniform per-surface scalars are passed as T.
template <bool CURV, class T, class S>
TL_HD AsphEval<T> asph_eval(const AsphSurfaceT<S> &s, T rho) {
  AsphEval<T> e;
  const T c(s.c);
  const T kc2((S(1) + s.k) * s.c * s.c);
  e.radicand = ffma(-kc2, rho, T(1));
  e.rs = frsqrt(e.radicand);
  const T root = e.radicand * e.rs;
  e.inv1r = frcp(T(1) + root);
  T p = T(s.a[6]), dp = T(S(8) * s.a[6]), ddp = T(S(56) * s.a[6]);
  const S m1[6] = {7, 6, 5, 4, 3, 2};
  const S m2[6] = {42, 30, 20, 12, 6, 2};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < 6; ++j) {
    const S coef = s.a[5 - j];
    p = ffma(p, rho, T(coef));
    dp = ffma(dp, rho, T(m1[j] * coef));
    if (CURV) ddp = ffma(ddp, rho, T(m2[j] * coef));
  }
  e.sag = ffma((c * rho), e.inv1r, (p * rho) * rho);
  e.slope = ffma(T(S(0.5) * s.c), e.rs, dp * rho);
  if (CURV) e.curv = ffma(T(S(0.25) * ((S(1) + s.k) * s.c * s.c) * s.c), (e.rs * e.rs) * e.rs, ddp);   // (uniform factor: scalar)
  else e.curv = T(0);
  return e;
}

// Early exit of the fast policy's Newton loop.  The oracle runs kNewton = 4 steps whatever happens; from the
// base-sphere start the iteration converges quadratically, e_{n+1} = M step_n^2 with M = F'' / (2 F') (M |tau| is
// 3e-4 ... 1e-2 on the config-3 lens), so after a step of at most TL_NEWTON_EXIT |tau| the steps still to come move
// tau by <= ~1e-8 |tau|, below fp32 resolution: they are skipped.  A thread leaves the loop when ALL its rays have
// settled (NaN / tau = 0 never do: those run the four steps); the exact policy keeps the fixed count.  Measured
// on the config-3 lens (fp32, 16 fields x 3 wavelengths x 48^2 pupil): every ray settles within two steps, the
// spherical surfaces of a mixed lens within one (skipping the loop there by a per-surface flag was slower).  B200:
// fused pass 1.005 -> 0.8105 ms, forward sweep 0.478 -> 0.295 ms at 4.2 M rays.  The fp64 instantiation (CPU gradient checks against the fp64
// oracle) uses a bound that keeps the skipped steps below 1e-12 |tau|.
#ifndef TL_NEWTON_EXIT
#define TL_NEWTON_EXIT 1e-3f
#endif
template <class T> TL_HD T newton_exit() { return T(TL_NEWTON_EXIT); }
template <> TL_HD double newton_exit<double>() { return 1e-6; }
template <class T>
TL_HD bool newton_settled(T, T) { return false; }                   // (lane types without an early exit)
TL_HD bool newton_settled(float step, float tau) { return fabsf(step) <= newton_exit<float>() * fabsf(tau); }
TL_HD bool newton_settled(double step, double tau) { return fabs(step) <= newton_exit<double>() * fabs(tau); }
// (`&`, not `&&`: one chain of predicated compares and ONE branch per step instead of a branch per ray)
TL_HD bool newton_settled(f2 step, f2 tau) {
  return newton_settled(step.v.x, tau.v.x) & newton_settled(step.v.y, tau.v.y);
}
TL_HD bool newton_settled(f4 step, f4 tau) { return newton_settled(step.a, tau.a) & newton_settled(step.b, tau.b); }
// ... and was that last step so small (<= 2e-6 |tau|; fp64: 1e-11) that the sag and slope just evaluated in front
// of it are those of the new point to working precision?  Then the evaluation at the hit point that follows the
// loop is skipped too: a surface costs as many evaluations as Newton steps (two on an aspheric surface of the
// config-3 lens, one on a spherical one) instead of one more.  Steps between the two bounds -- or rounding noise
// above the small one -- settle the loop and are followed by the evaluation as before: no cliff.
template <class T> TL_HD T newton_fresh() { return T(2e-6f); }
template <> TL_HD double newton_fresh<double>() { return 1e-11; }
template <class T>
TL_HD bool newton_is_fresh(T, T) { return false; }
TL_HD bool newton_is_fresh(float step, float tau) { return fabsf(step) <= newton_fresh<float>() * fabsf(tau); }
TL_HD bool newton_is_fresh(double step, double tau) { return fabs(step) <= newton_fresh<double>() * fabs(tau); }
TL_HD bool newton_is_fresh(f2 step, f2 tau) {
  return newton_is_fresh(step.v.x, tau.v.x) & newton_is_fresh(step.v.y, tau.v.y);
}
TL_HD bool newton_is_fresh(f4 step, f4 tau) { return newton_is_fresh(step.a, tau.a) & newton_is_fresh(step.b, tau.b); }

// One general surface, fast policy.  Returns through r (state behind the surface, z shifted),
// hit point in (hit_x, hit_y); tracks predicate margins like fast_surface.
template <class T, class S>
TL_HD void fast_asph_surface(Ray<T> &r, const AsphSurfaceT<S> &s, T &min_cos2, T &travel, T &min_clip,
                             T index, T &opl) {
  const T c(s.c), mu(s.mu);
  // base-sphere start
  const T ne = ffma(r.z, r.cz, ffma(r.y, r.cy, r.x * r.cx));
  const T mz = ffma(-ne, r.cz, r.z);
  const T m2 = ffma(-ne, ne, ffma(r.z, r.z, ffma(r.y, r.y, r.x * r.x)));
  const T tmp = ffma(c, m2, T(-2) * mz);
  const T q = ffma(-c, tmp, r.cz * r.cz);
  const T ci = q * frsqrt(q);
  T tau = ffma(tmp, frcp(r.cz + ci), -ne);
  AsphEval<T> e;
  bool fresh = false;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int it = 0; it < kNewton; ++it) {
    const T hx = ffma(tau, r.cx, r.x), hy = ffma(tau, r.cy, r.y);
    e = asph_eval<false>(s, ffma(hy, hy, hx * hx));
    const T f = ffma(tau, r.cz, r.z) - e.sag;
    const T fp = ffma(-(e.slope + e.slope), ffma(hy, r.cy, hx * r.cx), r.cz);
    const T rfp = frcp(fp);
    const T step = f * rfp;
    tau = ffma(-f, rfp, tau);
    if (newton_settled(step, tau)) {
      fresh = newton_is_fresh(step, tau);
      break;
    }
  }
  travel = tau * r.cz;
  r.x = ffma(tau, r.cx, r.x);
  r.y = ffma(tau, r.cy, r.y);
  r.z = r.z + travel;
  const T rho = ffma(r.y, r.y, r.x * r.x);
  if (!fresh) e = asph_eval<false>(s, rho);
  opl = ffma(index, tau, opl);
  min_clip = fmin2(min_clip, T(s.sd2) - rho);
  // unit normal, vector Snell
  const T mx = T(-2) * r.x * e.slope, my = T(-2) * r.y * e.slope;
  const T inv = frsqrt(ffma(my, my, ffma(mx, mx, T(1))));
  const T nx = mx * inv, ny = my * inv;
  const T cin = ffma(r.cz, inv, ffma(r.cy, ny, r.cx * nx));
  const T qo = ffma(-(mu * mu), ffma(-cin, cin, T(1)), T(1));
  const T co = qo * frsqrt(qo);
  const T g = ffma(-mu, cin, co);
  r.cx = ffma(g, nx, mu * r.cx);
  r.cy = ffma(g, ny, mu * r.cy);
  const T w = T(1) - ffma(r.cy, r.cy, r.cx * r.cx);
  r.cz = w * frsqrt(w);
  min_cos2 = fmin2(min_cos2, fmin2(fmin2(q, e.radicand), fmin2(qo, w)));
  r.z = r.z - T(s.t);
}

// ---------------------------------------------------------------------------
// Geometric adjoint of a general surface (see the derivation in trace_core.cuh):
//   F(h) = h_z - s(rho) = 0,  m = grad F = (-2 hx s', -2 hy s', 1),  n = m / |m|.
// Parameter gradients: g[0] = d/dc, g[1] = d/dk, g[2..8] = d/da4..a16.
// ---------------------------------------------------------------------------
template <class T>
struct AsphGrad {
  T p[kAsphParams];
  T t, mu;
};

// Seed on the optical path length (row A10): opl = sum_k n_k tau_k + n_S tau_image, n_0 = 1, n_{k+1} = n_k / mu_k.
// With q = d loss / d opl of this ray, tau_k gets the extra adjoint q n_k -- it joins gh.d, the adjoint the hit
// point induces on tau_k, in front of the implicit-function transfer, and from there reaches the surface
// parameters, the thickness in front, the ray's point and direction like any other seed -- and mu_k the extra
// gradient -q (sum_{j>k} n_j tau_j) / mu_k (every index behind surface k is proportional to 1 / mu_k).
// `tail` is that sum, carried along the backward walk.
template <class T>
struct OplSeed {
  T q, tail;
};
// ... at the image plane: tau_image = -z / cz gets q n_S next to (gx cx + gy cy)
template <class T>
TL_HD void sweep_begin_opl(Sweep<T> &s, const Ray<T> &pre, T q, T n_image) {
  s.gr.z = ffma(-(q * n_image), frcp(pre.cz), s.gr.z);
}

template <class T, class S>
TL_HD AsphGrad<T> sweep_asphere(Sweep<T> &s, T hx, T hy, T dx, T dy, const AsphSurfaceT<S> &sf,
                                OplSeed<T> *opl = nullptr, T n_in = T(0), T n_out = T(0)) {
  AsphGrad<T> g;
  const T c(sf.c), mu(sf.mu);
  const T rho = ffma(hy, hy, hx * hx);
  const AsphEval<T> e = asph_eval<true>(sf, rho);
  const T hz = e.sag;
  const T wd = ffma(-dy, dy, ffma(-dx, dx, T(1)));
  const T dz = wd * frsqrt(wd);
  const T dist = ffma((s.hit.z - hz) + T(sf.t), s.dir.z, ffma(s.hit.y - hy, s.dir.y, (s.hit.x - hx) * s.dir.x));
  const Vec3<T> gdo{ffma(dist, s.gr.x, s.gd.x), ffma(dist, s.gr.y, s.gd.y), ffma(dist, s.gr.z, s.gd.z)};
  g.t = -s.gr.z;
  // normal
  const T two_sp = e.slope + e.slope;
  const Vec3<T> m{-two_sp * hx, -two_sp * hy, T(1)};
  const T inv = frsqrt(ffma(m.y, m.y, ffma(m.x, m.x, T(1))));
  const Vec3<T> n{m.x * inv, m.y * inv, inv};
  const Vec3<T> d{dx, dy, dz};
  // refraction d' = mu d + g n
  const T a = dot3(n, d);
  const T ap = dot3(n, s.dir);
  const T gsn = ffma(-mu, a, ap);
  const T gdd = dot3(gdo, d);
  const T u = dot3(gdo, n) * frcp(ap);
  const T ga = -(mu * gsn) * u;
  g.mu = ffma(-u, ffma(a, ap, mu * ffma(-a, a, T(1))), gdd);
  if (opl) {                                             // (`dist` = tau of the segment behind this surface)
    opl->tail = ffma(n_out, dist, opl->tail);
    g.mu = ffma(-(opl->q * opl->tail), frcp(mu), g.mu);
  }
  const Vec3<T> gn{ffma(ga, d.x, gsn * gdo.x), ffma(ga, d.y, gsn * gdo.y), ffma(ga, d.z, gsn * gdo.z)};
  const Vec3<T> gdi{ffma(ga, n.x, mu * gdo.x), ffma(ga, n.y, mu * gdo.y), ffma(ga, n.z, mu * gdo.z)};
  // n = m / |m|  ->  gm = (gn - (gn.n) n) / |m|   (m_z is the constant 1)
  const T gnn = dot3(gn, n);
  const T gmx = ffma(-gnn, n.x, gn.x) * inv, gmy = ffma(-gnn, n.y, gn.y) * inv;
  const T qq = ffma(gmy, hy, gmx * hx);                  // d m / d s' contraction
  const T four_q_spp = T(4) * qq * e.curv;
  const Vec3<T> gh{ffma(-four_q_spp, hx, ffma(-two_sp, gmx, s.gr.x)),
                   ffma(-four_q_spp, hy, ffma(-two_sp, gmy, s.gr.y)), s.gr.z};
  // transfer onto the surface (implicit function theorem)
  const T gtau = opl ? ffma(opl->q, n_in, dot3(gh, d)) : dot3(gh, d);
  const T sd = gtau * frcp(dot3(m, d));
  s.gr = Vec3<T>{ffma(-sd, m.x, gh.x), ffma(-sd, m.y, gh.y), gh.z - sd};
  // parameters: g_theta = sd * ds/dtheta - 2 qq * ds'/dtheta
  const T m2q = T(-2) * qq;
  const T rs3 = (e.rs * e.rs) * e.rs;
  const S c3s = (sf.c * sf.c) * sf.c;                    // (uniform over the warp: scalar arithmetic, broadcast operands)
  const T c3q(S(0.25) * c3s), c3h(S(0.5) * c3s);
  g.p[0] = ffma(m2q * T(S(0.5)), rs3, sd * ((rho * e.rs) * e.inv1r));
  g.p[1] = ffma(m2q, (c3q * rho) * rs3, sd * (c3h * (rho * rho) * e.rs * (e.inv1r * e.inv1r)));
  // a_i (i = 2..8): sd rho^i - 2 qq i rho^(i-1) = rho^(i-1) (sd rho + i m2q)
  const T sd_rho = sd * rho;
  T pw = rho;                                            // rho^(i-1)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 2; i <= 8; ++i) {
    g.p[i] = pw * ffma(m2q, T(S(i)), sd_rho);
    pw = pw * rho;
  }
  s.gd = gdi;
  s.hit = Vec3<T>{hx, hy, hz};
  s.dir = d;
  return g;
}

}  // namespace tl
