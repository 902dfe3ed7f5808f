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
                                  T *ox, T *oy, T *ocx, T *ocy, T *oopl, T *min_cos2, T *min_clip,
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
    for (int k = S - 1; k >= 0; --k) {
      const AsphSurfaceT<T> sf = make_surface<T>(k, c, kk, a, t, mu, sd2);
      AsphGrad<T> g = sweep_asphere(sw, st[k + 1].x, st[k + 1].y, st[k].cx, st[k].cy, sf);
      for (int j = 0; j < kAsphParams; ++j) gp[k * kAsphParams + j] += (double)g.p[j];
      gt[k] += (double)g.t; gmu[k] += (double)g.mu;
    }
    sweep_end(sw, z[i], gx[i], gy[i], gz[i], gcx[i], gcy[i]);
  }
}

#define HCA_ARGS(T)                                                                              \
  int64_t n, const T *x, const T *y, const T *z, const T *cx, const T *cy, int S, const T *c,   \
      const T *kk, const T *a, const T *t, const T *mu, const T *sd2, const T *sx, const T *sy, \
      const T *scx, const T *scy, T *ox, T *oy, T *ocx, T *ocy, T *oopl, T *min_cos2,           \
      T *min_clip, T *gx, T *gy, T *gz, T *gcx, T *gcy, double *gp, double *gt, double *gmu
#define HCA_PASS n, x, y, z, cx, cy, S, c, kk, a, t, mu, sd2, sx, sy, scx, scy, ox, oy, ocx, ocy, oopl, \
                 min_cos2, min_clip, gx, gy, gz, gcx, gcy, gp, gt, gmu
extern "C" void hc_asph_fast_f32(HCA_ARGS(float)) { asph_fast_and_adjoint<float>(HCA_PASS); }
extern "C" void hc_asph_fast_f64(HCA_ARGS(double)) { asph_fast_and_adjoint<double>(HCA_PASS); }
