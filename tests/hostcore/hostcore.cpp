// TEST AID ONLY: runs the per-ray arithmetic of torchoptics_b200/csrc/trace_core.cuh
// on the CPU so that tests/ can check it against the oracle without a GPU
// (exact policy: bit-for-bit; fast policy and the adjoint: fp64 against autograd).
// Built by tests/hostcore/build.py with g++ -ffp-contract=off.  The product never
// loads this library.
#include <stdint.h>
#include <vector>
#include "../../torchoptics_b200/csrc/trace_core.cuh"
#include "../../torchoptics_b200/csrc/trace_core_asph.cuh"

using namespace tl;

extern "C" {

// one lens, one wavelength; rays are [n]
void hc_trace_exact(int64_t n, const float *x, const float *y, const float *z, const float *cx,
                    const float *cy, int S, const float *c, const float *t, const float *mu,
                    const uint8_t *live, int allow_backward, float *ox, float *oy, float *ocx,
                    float *ocy, uint8_t *ook, uint8_t *obw) {
  for (int64_t i = 0; i < n; ++i) {
    Ray<float> r{x[i], y[i], z[i], cx[i], cy[i], exact_cz0(cx[i], cy[i])};
    bool ok = true, bw = false;
    for (int k = 0; k < S; ++k) {
      Surface s{c[k], t[k], mu[k]};
      exact_surface(r, s, k > 0 && live[k - 1], allow_backward != 0, ok, bw);
    }
    exact_image(r, live[S - 1] != 0, allow_backward != 0, ok, bw);
    ox[i] = r.x; oy[i] = r.y; ocx[i] = r.cx; ocy[i] = r.cy; ook[i] = ok; obw[i] = bw;
  }
}

}  // extern "C"

template <class T>
static void fast_and_adjoint(int64_t n, const T *x, const T *y, const T *z, const T *cx, const T *cy,
                             int S, const T *c, const T *t, const T *mu, const uint8_t *live,
                             const T *sx, const T *sy, const T *scx, const T *scy,
                             T *ox, T *oy, T *ocx, T *ocy, T *min_cos2, T *min_travel,
                             T *gx, T *gy, T *gz, T *gcx, T *gcy, double *gc, double *gt,
                             double *gmu) {
  std::vector<Ray<T>> st(S + 1);
  for (int k = 0; k < S; ++k) gc[k] = gt[k] = gmu[k] = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    Ray<T> r{x[i], y[i], z[i], cx[i], cy[i], fast_cz0(cx[i], cy[i])};
    T mq = T(1), mtr = T(1e30);
    for (int k = 0; k < S; ++k) {
      st[k] = r;
      T travel;
      fast_surface(r, c[k], mu[k], mu[k] * mu[k], t[k], mq, travel);
      if (k > 0 && live[k - 1]) mtr = fmin2(mtr, travel);
    }
    st[S] = r;
    Ray<T> pre = r;
    T travel = fast_image(r);
    if (live[S - 1]) mtr = fmin2(mtr, travel);
    ox[i] = r.x; oy[i] = r.y; ocx[i] = r.cx; ocy[i] = r.cy;
    min_cos2[i] = mq; min_travel[i] = mtr;
    if (!sx) continue;
    // like the kernels: geometric sweep over the parked (hit x, hit y, in-dir x, in-dir y)
    Sweep<T> sw = sweep_begin(pre, r.x, r.y, sx[i], sy[i], scx[i], scy[i]);
    for (int k = S - 1; k >= 0; --k) {
      SurfaceGrad<T> g = sweep_sphere(sw, st[k + 1].x, st[k + 1].y, st[k].cx, st[k].cy, c[k], t[k], mu[k],
                                      mu[k] * mu[k]);
      gc[k] += (double)g.c; gt[k] += (double)g.t; gmu[k] += (double)g.mu;
    }
    Ray<T> a;
    sweep_end(sw, z[i], a.x, a.y, a.z, a.cx, a.cy);
    gx[i] = a.x; gy[i] = a.y; gz[i] = a.z; gcx[i] = a.cx; gcy[i] = a.cy;
  }
}

extern "C" {
#define HC_ARGS(T)                                                                              \
  int64_t n, const T *x, const T *y, const T *z, const T *cx, const T *cy, int S, const T *c,  \
      const T *t, const T *mu, const uint8_t *live, const T *sx, const T *sy, const T *scx,    \
      const T *scy, T *ox, T *oy, T *ocx, T *ocy, T *min_cos2, T *min_travel, T *gx, T *gy,     \
      T *gz, T *gcx, T *gcy, double *gc, double *gt, double *gmu
#define HC_PASS n, x, y, z, cx, cy, S, c, t, mu, live, sx, sy, scx, scy, ox, oy, ocx, ocy, \
                min_cos2, min_travel, gx, gy, gz, gcx, gcy, gc, gt, gmu

void hc_fast_f32(HC_ARGS(float)) { fast_and_adjoint<float>(HC_PASS); }
void hc_fast_f64(HC_ARGS(double)) { fast_and_adjoint<double>(HC_PASS); }

}  // extern "C"


// Reversible formulation (fast_surface_rev / sweep_sphere_rev): the forward parks (dist, cos, cos')
// per surface and the sweep walks the ray backwards.  EXACT_PARK: the parked values come from the
// exact-policy trace instead (what the kernels do for a lane that was re-traced).
template <class T>
static void rev_and_adjoint(int64_t n, const T *x, const T *y, const T *z, const T *cx, const T *cy,
                            int S, const T *c, const T *t, const T *mu, const uint8_t *live,
                            const T *sx, const T *sy, const T *scx, const T *scy,
                            T *ox, T *oy, T *ocx, T *ocy, T *min_cos2, T *min_travel,
                            T *gx, T *gy, T *gz, T *gcx, T *gcy, double *gc, double *gt,
                            double *gmu, int exact_park) {
  std::vector<Parked<T>> st(S);
  for (int k = 0; k < S; ++k) gc[k] = gt[k] = gmu[k] = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    Ray<T> r{x[i], y[i], z[i], cx[i], cy[i], fast_cz0(cx[i], cy[i])};
    T mq = T(1), mcz = T(1), mtr = T(1e30);
    for (int k = 0; k < S; ++k) {
      T travel;
      fast_surface_rev(r, c[k], mu[k], mu[k] * mu[k], T(1) - mu[k] * mu[k], t[k], mq, mcz, travel, st[k]);
      if (k > 0 && live[k - 1]) mtr = fmin2(mtr, travel);
    }
    Ray<T> pre = r;
    T travel = fast_image(r);
    if (live[S - 1]) mtr = fmin2(mtr, travel);
    if (exact_park & 1) {      // float only: re-trace with the exact policy, park its values
      Ray<float> e{(float)x[i], (float)y[i], (float)z[i], (float)cx[i], (float)cy[i],
                   exact_cz0((float)cx[i], (float)cy[i])};
      bool ok = true, bw = false;
      for (int k = 0; k < S; ++k) {
        Surface sf{(float)c[k], (float)t[k], (float)mu[k]};
        Parked<float> pk;
        exact_surface_t<false>(e, sf, k > 0 && live[k - 1], true, ok, bw, nullptr, &pk);
        st[k].dist = pk.dist; st[k].ci = pk.ci; st[k].co = pk.co;
      }
      pre = Ray<T>{T(e.x), T(e.y), T(e.z), T(e.cx), T(e.cy), T(e.cz)};
      exact_image(e, live[S - 1] != 0, true, ok, bw);
      r.x = e.x; r.y = e.y;
    }
    ox[i] = r.x; oy[i] = r.y; ocx[i] = pre.cx; ocy[i] = pre.cy;
    min_cos2[i] = fmin2(mq, mcz > T(kBandCz) ? T(1) : T(0)); min_travel[i] = mtr;
    if (!sx) continue;
    SweepRev<T> sw = sweep_begin_rev(pre, r.x, r.y, sx[i], sy[i], scx[i], scy[i]);
    for (int k = S - 1; k >= 0; --k) {
      SurfaceGrad<T> g = (exact_park & 2)
          ? sweep_sphere_rev2(sw, st[k].dist, st[k].ci, c[k], t[k], mu[k], mu[k] * mu[k], T(1) - mu[k] * mu[k],
                              T(1) / mu[k])
          : sweep_sphere_rev(sw, st[k], c[k], t[k], mu[k], T(1) / mu[k]);
      gc[k] += (double)g.c; gt[k] += (double)g.t; gmu[k] += (double)g.mu;
    }
    sweep_end_rev(sw, gx[i], gy[i], gz[i], gcx[i], gcy[i]);
  }
}

extern "C" {
void hc_rev_f32(HC_ARGS(float), int exact_park) { rev_and_adjoint<float>(HC_PASS, exact_park); }
void hc_rev_f64(HC_ARGS(double), int exact_park) { rev_and_adjoint<double>(HC_PASS, exact_park); }
}

// ---------------------------------------------------------------------------
// extension surfaces (one lens, one wavelength); surface tables: c,k,t,mu,sd2 [S], a [S,7]
// ---------------------------------------------------------------------------
template <class S>
static AsphSurfaceT<S> make_surface(int k, const S *c, const S *kk, const S *a, const S *t, const S *mu,
                                    const S *sd2) {
  AsphSurfaceT<S> s;
  s.c = c[k]; s.k = kk[k]; s.t = t[k]; s.mu = mu[k]; s.sd2 = sd2[k];
  for (int j = 0; j < kAsphCoefs; ++j) s.a[j] = a[k * kAsphCoefs + j];
  return s;
}

extern "C" void hc_asph_exact(int64_t n, const float *x, const float *y, const float *z, const float *cx,
                              const float *cy, int S, const float *c, const float *kk, const float *a,
                              const float *t, const float *mu, const float *sd2, const uint8_t *live,
                              int allow_backward, float *ox, float *oy, float *ocx, float *ocy,
                              uint8_t *ook, uint8_t *obw, float *oopl) {
  for (int64_t i = 0; i < n; ++i) {
    Ray<float> r{x[i], y[i], z[i], cx[i], cy[i], exact_cz0(cx[i], cy[i])};
    bool ok = true, bw = false;
    float index = 1.0f, opl = 0.0f;
    for (int k = 0; k < S; ++k)
      exact_asph_surface(r, make_surface<float>(k, c, kk, a, t, mu, sd2), k > 0 && live[k - 1],
                         allow_backward != 0, ok, bw, index, opl);
    exact_asph_image(r, live[S - 1] != 0, allow_backward != 0, ok, bw, index, opl);
    ox[i] = r.x; oy[i] = r.y; ocx[i] = r.cx; ocy[i] = r.cy; ook[i] = ok; obw[i] = bw; oopl[i] = opl;
  }
}

template <class T>
static void asph_fast_and_adjoint(int64_t n, const T *x, const T *y, const T *z, const T *cx, const T *cy,
                                  int S, const T *c, const T *kk, const T *a, const T *t, const T *mu,
                                  const T *sd2, const T *sx, const T *sy, const T *scx, const T *scy,
                                  const T *sopl, T *ox, T *oy, T *ocx, T *ocy, T *oopl, T *min_cos2, T *min_clip,
                                  T *gx, T *gy, T *gz, T *gcx, T *gcy, double *gp, double *gt,
                                  double *gmu) {
  std::vector<Ray<T>> st(S + 1);
  for (int k = 0; k < S; ++k) {
    gt[k] = gmu[k] = 0.0;
    for (int j = 0; j < kAsphParams; ++j) gp[k * kAsphParams + j] = 0.0;
  }
  for (int64_t i = 0; i < n; ++i) {
    Ray<T> r{x[i], y[i], z[i], cx[i], cy[i], fast_cz0(cx[i], cy[i])};
    T mq = T(1), mclip = T(1e30), index = T(1), opl = T(0);
    for (int k = 0; k < S; ++k) {
      st[k] = r;
      T travel;
      const AsphSurfaceT<T> sf = make_surface<T>(k, c, kk, a, t, mu, sd2);
      fast_asph_surface(r, sf, mq, travel, mclip, index, opl);
      index = index / mu[k];
    }
    st[S] = r;
    Ray<T> pre = r;
    fast_image(r);
    opl = opl + index * (-pre.z / pre.cz);
    ox[i] = r.x; oy[i] = r.y; ocx[i] = r.cx; ocy[i] = r.cy; oopl[i] = opl;
    min_cos2[i] = mq; min_clip[i] = mclip;
    if (!sx) continue;
    Sweep<T> sw = sweep_begin(pre, r.x, r.y, sx[i], sy[i], scx[i], scy[i]);
    // (the index in front of every surface, as the kernel's table holds it)
    std::vector<T> n_at(S + 1);
    n_at[0] = T(1);
    for (int k = 0; k < S; ++k) n_at[k + 1] = n_at[k] / mu[k];
    OplSeed<T> os{sopl ? sopl[i] : T(0), T(0)};
    if (sopl) sweep_begin_opl(sw, pre, os.q, n_at[S]);
    for (int k = S - 1; k >= 0; --k) {
      const AsphSurfaceT<T> sf = make_surface<T>(k, c, kk, a, t, mu, sd2);
      AsphGrad<T> g = sweep_asphere(sw, st[k + 1].x, st[k + 1].y, st[k].cx, st[k].cy, sf, sopl ? &os : nullptr,
                                    n_at[k], n_at[k + 1]);
      for (int j = 0; j < kAsphParams; ++j) gp[k * kAsphParams + j] += (double)g.p[j];
      gt[k] += (double)g.t; gmu[k] += (double)g.mu;
    }
    sweep_end(sw, z[i], gx[i], gy[i], gz[i], gcx[i], gcy[i]);
  }
}

#define HCA_ARGS(T)                                                                              \
  int64_t n, const T *x, const T *y, const T *z, const T *cx, const T *cy, int S, const T *c,   \
      const T *kk, const T *a, const T *t, const T *mu, const T *sd2, const T *sx, const T *sy, \
      const T *scx, const T *scy, const T *sopl, T *ox, T *oy, T *ocx, T *ocy, T *oopl, T *min_cos2, \
      T *min_clip, T *gx, T *gy, T *gz, T *gcx, T *gcy, double *gp, double *gt, double *gmu
#define HCA_PASS n, x, y, z, cx, cy, S, c, kk, a, t, mu, sd2, sx, sy, scx, scy, sopl, ox, oy, ocx, ocy, oopl, \
                 min_cos2, min_clip, gx, gy, gz, gcx, gcy, gp, gt, gmu
extern "C" void hc_asph_fast_f32(HCA_ARGS(float)) { asph_fast_and_adjoint<float>(HCA_PASS); }
extern "C" void hc_asph_fast_f64(HCA_ARGS(double)) { asph_fast_and_adjoint<double>(HCA_PASS); }


// ---------------------------------------------------------------------------
// aggregate=True penalty terms (rtl:641-657): exact policy, fast policy and the adjoint with
// per-surface seeds.  Stacks are [S, n] (surface-major).
// ---------------------------------------------------------------------------
extern "C" void hc_trace_exact_pen(int64_t n, const float *x, const float *y, const float *z,
                                   const float *cx, const float *cy, int S, const float *c,
                                   const float *t, const float *mu, const uint8_t *live,
                                   int allow_backward, float *zr, float *th, float *thp, uint8_t *ook,
                                   uint32_t *okbits) {
  for (int64_t i = 0; i < n; ++i) {
    Ray<float> r{x[i], y[i], z[i], cx[i], cy[i], exact_cz0(cx[i], cy[i])};
    bool ok = true, bw = false;
    uint32_t bits = 0;
    for (int k = 0; k < S; ++k) {
      Surface s{c[k], t[k], mu[k]};
      Penalty pen;
      exact_surface_t<true>(r, s, k > 0 && live[k - 1], allow_backward != 0, ok, bw, &pen);
      zr[k * n + i] = pen.z_relu; th[k * n + i] = pen.theta; thp[k * n + i] = pen.theta_prime;
      if (ok) bits |= 1u << k;
    }
    exact_image(r, live[S - 1] != 0, allow_backward != 0, ok, bw);
    ook[i] = ok;
    okbits[i] = bits;
  }
}

template <class T>
static void fast_pen_and_adjoint(int64_t n, const T *x, const T *y, const T *z, const T *cx, const T *cy,
                                 int S, const T *c, const T *t, const T *mu, const T *sy,
                                 const T *szr, const T *sth, const T *sthp, T *zr, T *th, T *thp,
                                 T *gx, T *gy, T *gz, T *gcx, T *gcy, double *gc, double *gt,
                                 double *gmu) {
  std::vector<Ray<T>> st(S + 1);
  for (int k = 0; k < S; ++k) gc[k] = gt[k] = gmu[k] = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    Ray<T> r{x[i], y[i], z[i], cx[i], cy[i], fast_cz0(cx[i], cy[i])};
    T mq = T(1);
    for (int k = 0; k < S; ++k) {
      st[k] = r;
      T travel, ci, co;
      fast_surface(r, c[k], mu[k], mu[k] * mu[k], t[k], mq, travel, ci, co);
      zr[k * n + i] = r.z > T(0) ? r.z : T(0);
      th[k * n + i] = fast_angle_norm(ci);
      thp[k * n + i] = fast_angle_norm(co);
    }
    st[S] = r;
    Ray<T> pre = r;
    fast_image(r);
    Sweep<T> sw = sweep_begin(pre, r.x, r.y, T(0), sy[i], T(0), T(0));
    for (int k = S - 1; k >= 0; --k) {
      SurfaceGrad<T> g = sweep_sphere_pen(sw, st[k + 1].x, st[k + 1].y, st[k].cx, st[k].cy, c[k], t[k],
                                          mu[k], mu[k] * mu[k], szr[k * n + i], sth[k * n + i],
                                          sthp[k * n + i], T(1));
      gc[k] += (double)g.c; gt[k] += (double)g.t; gmu[k] += (double)g.mu;
    }
    sweep_end(sw, z[i], gx[i], gy[i], gz[i], gcx[i], gcy[i]);
  }
}

#define HCP_ARGS(T)                                                                             \
  int64_t n, const T *x, const T *y, const T *z, const T *cx, const T *cy, int S, const T *c,  \
      const T *t, const T *mu, const T *sy, const T *szr, const T *sth, const T *sthp, T *zr,  \
      T *th, T *thp, T *gx, T *gy, T *gz, T *gcx, T *gcy, double *gc, double *gt, double *gmu
#define HCP_PASS n, x, y, z, cx, cy, S, c, t, mu, sy, szr, sth, sthp, zr, th, thp, gx, gy, gz, gcx, gcy, \
                 gc, gt, gmu
extern "C" void hc_fast_pen_f32(HCP_ARGS(float)) { fast_pen_and_adjoint<float>(HCP_PASS); }
extern "C" void hc_fast_pen_f64(HCP_ARGS(double)) { fast_pen_and_adjoint<double>(HCP_PASS); }

// The penalty backward exactly as k_trace_adj<.., MODE_BWD, .., PEN> runs it for a ray that took
// the exact policy: exact trace with parked (hit, incoming direction) and ok bits, then the sweep
// with seeds only where the ray is ok behind the surface and every other lane state forced to zero.
template <class T>
static void exact_pen_adjoint(int64_t n, const float *x, const float *y, const float *z, const float *cx,
                              const float *cy, int S, const float *c, const float *t, const float *mu,
                              const uint8_t *live, int allow_backward, const T *sy, const T *szr,
                              const T *sth, const T *sthp, T *gx, T *gy, T *gz, T *gcx, T *gcy,
                              double *gc, double *gt, double *gmu) {
  std::vector<float> hx(S), hy(S), dx(S), dy(S);
  for (int k = 0; k < S; ++k) gc[k] = gt[k] = gmu[k] = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    Ray<float> r{x[i], y[i], z[i], cx[i], cy[i], exact_cz0(cx[i], cy[i])};
    bool ok = true, bw = false;
    uint32_t bits = 0, flips = 0;
    for (int k = 0; k < S; ++k) {
      const float in_cx = r.cx, in_cy = r.cy;
      Surface s{c[k], t[k], mu[k]};
      exact_surface(r, s, k > 0 && live[k - 1], allow_backward != 0, ok, bw);
      hx[k] = r.x; hy[k] = r.y; dx[k] = in_cx; dy[k] = in_cy;
      if (ok) bits |= 1u << k;
      if (beyond_equator(c[k], r.z, t[k])) flips |= 1u << k;
    }
    Ray<T> pre{T(r.x), T(r.y), T(r.z), T(r.cx), T(r.cy), T(r.cz)};
    exact_image(r, live[S - 1] != 0, allow_backward != 0, ok, bw);
    Sweep<T> sw = sweep_begin(pre, T(r.x), T(r.y), T(0), ok ? sy[i] : T(0), T(0), T(0));
    for (int k = S - 1; k >= 0; --k) {
      const bool ok_k = (bits >> k) & 1u;
      SurfaceGrad<T> g = sweep_sphere_pen(sw, T(hx[k]), T(hy[k]), T(dx[k]), T(dy[k]), T(c[k]), T(t[k]),
                                          T(mu[k]), T(mu[k]) * T(mu[k]), ok_k ? szr[k * n + i] : T(0),
                                          ok_k ? sth[k * n + i] : T(0), ok_k ? sthp[k * n + i] : T(0),
                                          ((flips >> k) & 1u) ? T(-1) : T(1));
      if (!ok_k) {
        g.c = T(0); g.mu = T(0);
        g.t = (-t[k] > 0.f) ? -szr[k * n + i] : T(0);
        sw.gr = Vec3<T>{T(0), T(0), T(0)};
        sw.gd = Vec3<T>{T(0), T(0), T(0)};
      }
      gc[k] += (double)g.c; gt[k] += (double)g.t; gmu[k] += (double)g.mu;
    }
    sweep_end(sw, T(z[i]), gx[i], gy[i], gz[i], gcx[i], gcy[i]);
  }
}

#define HCE_ARGS(T)                                                                                  \
  int64_t n, const float *x, const float *y, const float *z, const float *cx, const float *cy, int S, \
      const float *c, const float *t, const float *mu, const uint8_t *live, int allow_backward,       \
      const T *sy, const T *szr, const T *sth, const T *sthp, T *gx, T *gy, T *gz, T *gcx, T *gcy,   \
      double *gc, double *gt, double *gmu
#define HCE_PASS n, x, y, z, cx, cy, S, c, t, mu, live, allow_backward, sy, szr, sth, sthp, gx, gy, gz, \
                 gcx, gcy, gc, gt, gmu
extern "C" void hc_exact_pen_adjoint_f32(HCE_ARGS(float)) { exact_pen_adjoint<float>(HCE_PASS); }
extern "C" void hc_exact_pen_adjoint_f64(HCE_ARGS(double)) { exact_pen_adjoint<double>(HCE_PASS); }

// Forward-mode pair D2 through the fast policy (what k_aim runs for its tee rays): image point
// and its Jacobian w.r.t. the entrance-pupil point, one lens, one wavelength.
extern "C" void hc_forward_mode(int64_t n, const float *x, const float *y, const float *z, const float *cx,
                                const float *cy, int S, const float *c, const float *t, const float *mu,
                                float *ox, float *oy, float *jac /* [n,4]: dx/dxp, dx/dyp, dy/dxp, dy/dyp */) {
  for (int64_t i = 0; i < n; ++i) {
    Ray<D2> r{D2(x[i], 1.f, 0.f), D2(y[i], 0.f, 1.f), D2(z[i]), D2(cx[i]), D2(cy[i]),
              fast_cz0(D2(cx[i]), D2(cy[i]))};
    D2 mq(1.0f), travel;
    for (int k = 0; k < S; ++k) fast_surface(r, D2(c[k]), D2(mu[k]), D2(mu[k] * mu[k]), D2(t[k]), mq, travel);
    fast_image(r);
    ox[i] = r.x.v; oy[i] = r.y.v;
    jac[4 * i + 0] = r.x.a; jac[4 * i + 1] = r.x.b; jac[4 * i + 2] = r.y.a; jac[4 * i + 3] = r.y.b;
  }
}

// ---- paraxial front end (csrc/paraxial.cuh): the per-lens functions the kernels k_paraxial_fwd / _bwd call ----
#include "../../include/torchoptics_b200.h"
#include "../../torchoptics_b200/csrc/paraxial.cuh"
extern "C" void hc_paraxial(int B, int L, int mode, const float *c, const float *t, const float *n,
                            const uint8_t *live, const uint8_t *glass, const float *gout, float *out, float *gc,
                            float *gt, float *gn) {
  TlParaxial p{c, t, n, live, glass, B, L, mode};
  for (int b = 0; b < B; ++b) {
    paraxial_fwd_one(p, b, out);
    if (gout) paraxial_bwd_one(p, b, gout, gc, gt, gn);
  }
}
