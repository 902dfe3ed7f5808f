// Paraxial (ABCD) front end of the batched-lens workload, SURVEY.md section 8(f)-3: the per-lens
// scalar work the reference does with ~30 eager tensor ops per call --
//
//   get_first_order          rtl:772-794   (EFL, BFL) of every lens
//   compute_last_curvature   rtl:725-769   the last free curvature solved so that EFL = 1
//   (interface_propagation_abcd rtl:314-327, reduce_abcd rtl:301-311 underneath)
//
// -- and that `Optical_Loss.optical_loss_unsupervised_single` (optical_loss.py:63-64) runs once per
// sample in a Python loop.  Here: ONE thread per lens walks its <= 64 padded slots, forward in one
// launch, the hand-derived adjoint (prefix products kept in local memory, reverse walk) in another.
// Included by trace_kernels.cu inside its anonymous namespace; the per-lens functions also compile
// for the host (tests/hostcore: the adjoint against autograd of the reference's formulas, no GPU).
//
// System matrix: M = M_{K-1} ... M_1 M_0 over the included slots, M_k = [[1 + t C, t D], [C, D]] with
// D = n_k / n_{k+1}, C = c (D - 1) (refract at curvature c, then travel t).  The reference multiplies
// pairwise in fp32 (log-depth tree); a thread here multiplies left to right in fp64 and rounds the
// results once -- closer to the exact value than either fp32 order, and within 1e-6 of the
// reference's (tests/test_gpu_paraxial.py against the reference-generated goldens).

#ifdef __CUDACC__
#define PX_HD __host__ __device__ __forceinline__
#else
#define PX_HD inline
#endif

struct ParaxialSlots {       // which slots enter the product, per lens
  int n_incl;                // slots [0, n_incl) that are live
  int zero_t_at;             // slot whose thickness counts as 0 (-1: none)
  int solve_at;              // TL_PARAXIAL_LAST_CURVATURE: slot of the solved curvature
};

PX_HD ParaxialSlots paraxial_slots(const TlParaxial &p, int b) {
  const uint8_t *live = p.live + (int64_t)b * p.L;
  const uint8_t *glass = p.glass + (int64_t)b * p.L;
  int n_surf = 0;
  for (int k = 0; k < p.L; ++k) n_surf += live[k] != 0;      // (prefix masks, like the reference: rtl:733, :780)
  ParaxialSlots s;
  if (p.mode == TL_PARAXIAL_FIRST_ORDER) {
    s.n_incl = n_surf;
    s.zero_t_at = n_surf - 1;                                 // "zero-out the last thickness" rtl:781-783
    s.solve_at = -1;
  } else {
    const bool air_air = n_surf >= 2 && glass[n_surf - 2] == 0;      // rtl:735-736
    s.solve_at = n_surf - 1 - (air_air ? 1 : 0);              // rtl:738
    s.n_incl = s.solve_at < 0 ? 0 : s.solve_at;               // neither the last surface nor the solved one: rtl:741, :750
    s.zero_t_at = -1;
  }
  return s;
}

struct PxAbcd {
  double a, b, c, d;
};

PX_HD PxAbcd paraxial_step(const TlParaxial &p, int b, int k, const ParaxialSlots &s, double &cap,
                                              double &ratio, double &tk) {
  const int64_t o = (int64_t)b * p.L + k;
  const double n_before = k > 0 ? (double)p.n[o - 1] : 1.0;
  const double n_after = (double)p.n[o];
  ratio = n_before / n_after;
  cap = (double)p.c[o] * (ratio - 1.0);
  tk = k == s.zero_t_at ? 0.0 : (double)p.t[o];
  return PxAbcd{1.0 + tk * cap, tk * ratio, cap, ratio};
}

// out [B,2]: FIRST_ORDER (efl, bfl); LAST_CURVATURE (solved curvature, its slot index)
PX_HD void paraxial_fwd_one(const TlParaxial &p, int b, float *out) {
  const ParaxialSlots s = paraxial_slots(p, b);
  PxAbcd m{1.0, 0.0, 0.0, 1.0};
  for (int k = 0; k < s.n_incl; ++k) {
    if (!p.live[(int64_t)b * p.L + k]) continue;
    double cap, ratio, tk;
    const PxAbcd e = paraxial_step(p, b, k, s, cap, ratio, tk);
    m = PxAbcd{e.a * m.a + e.b * m.c, e.a * m.b + e.b * m.d, e.c * m.a + e.d * m.c, e.c * m.b + e.d * m.d};
  }
  if (p.mode == TL_PARAXIAL_FIRST_ORDER) {
    out[2 * b] = (float)(-1.0 / m.c);              // EFL = -1 / C   rtl:788
    out[2 * b + 1] = (float)(-m.a / m.c);          // BFL = -A / C   rtl:791
  } else {
    const double n_last = s.solve_at > 0 ? (double)p.n[(int64_t)b * p.L + s.solve_at - 1] : 1.0;
    out[2 * b] = (float)(-(1.0 + n_last * m.c) / (m.a * (n_last - 1.0)));      // rtl:762
    out[2 * b + 1] = (float)s.solve_at;
  }
}

// gout [B,2] (the second column of LAST_CURVATURE is an index: ignored) -> gc, gt, gn [B,L]
PX_HD void paraxial_bwd_one(const TlParaxial &p, int b, const float *gout, float *gc, float *gt, float *gn) {
  const ParaxialSlots s = paraxial_slots(p, b);
  PxAbcd prefix[TL_PARAXIAL_MAX_SLOTS];      // product of the slots in front of k (local memory: 2 KB)
  double gn_acc[TL_PARAXIAL_MAX_SLOTS];    // d loss / d n_after[k]: a slot's index enters its own matrix and the next one's
  PxAbcd m{1.0, 0.0, 0.0, 1.0};
  for (int k = 0; k < p.L; ++k) {
    const int64_t o = (int64_t)b * p.L + k;
    gc[o] = gt[o] = 0.f;
    gn_acc[k] = 0.0;
  }
  for (int k = 0; k < s.n_incl; ++k) {
    prefix[k] = m;
    if (!p.live[(int64_t)b * p.L + k]) continue;
    double cap, ratio, tk;
    const PxAbcd e = paraxial_step(p, b, k, s, cap, ratio, tk);
    m = PxAbcd{e.a * m.a + e.b * m.c, e.a * m.b + e.b * m.d, e.c * m.a + e.d * m.c, e.c * m.b + e.d * m.d};
  }
  // seed: d loss / d (A, B, C, D) of the product
  PxAbcd h{0.0, 0.0, 0.0, 0.0};
  if (p.mode == TL_PARAXIAL_FIRST_ORDER) {
    const double g_efl = gout[2 * b], g_bfl = gout[2 * b + 1];
    h.c = (g_efl + g_bfl * m.a) / (m.c * m.c);      // efl = -1 / C, bfl = -A / C
    h.a = -g_bfl / m.c;
  } else {
    const double g = gout[2 * b];
    const double n_last = s.solve_at > 0 ? (double)p.n[(int64_t)b * p.L + s.solve_at - 1] : 1.0;
    const double solved = -(1.0 + n_last * m.c) / (m.a * (n_last - 1.0));
    h.c = -g * n_last / (m.a * (n_last - 1.0));
    h.a = -g * solved / m.a;
    if (s.solve_at > 0) gn_acc[s.solve_at - 1] = g * (1.0 + m.c) / (m.a * (n_last - 1.0) * (n_last - 1.0));
  }
  for (int k = s.n_incl - 1; k >= 0; --k) {
    const int64_t o = (int64_t)b * p.L + k;
    if (!p.live[o]) continue;
    double cap, ratio, tk;
    const PxAbcd e = paraxial_step(p, b, k, s, cap, ratio, tk);
    const PxAbcd q = prefix[k];
    // d M_k = H P_k^T, then H <- M_k^T H
    const double m00 = h.a * q.a + h.b * q.b, m01 = h.a * q.c + h.b * q.d;
    const double m10 = h.c * q.a + h.d * q.b, m11 = h.c * q.c + h.d * q.d;
    h = PxAbcd{e.a * h.a + e.c * h.c, e.a * h.b + e.c * h.d, e.b * h.a + e.d * h.c, e.b * h.b + e.d * h.d};
    const double g_cap = m00 * tk + m10;                                  // M_k = [[1 + t C, t D], [C, D]]
    const double g_ratio = m01 * tk + m11 + g_cap * (double)p.c[o];       // C = c (D - 1)
    if (k != s.zero_t_at) gt[o] = (float)(m00 * cap + m01 * ratio);
    gc[o] = (float)(g_cap * (ratio - 1.0));
    const double n_after = (double)p.n[o];                                // D = n_before / n_after
    gn_acc[k] -= g_ratio * ratio / n_after;
    if (k > 0) gn_acc[k - 1] += g_ratio / n_after;
  }
  for (int k = 0; k < p.L; ++k) gn[(int64_t)b * p.L + k] = (float)gn_acc[k];
}

#ifdef __CUDACC__
__global__ void k_paraxial_fwd(TlParaxial p, float *out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < p.B) paraxial_fwd_one(p, b, out);
}

__global__ void k_paraxial_bwd(TlParaxial p, const float *gout, float *gc, float *gt, float *gn) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < p.B) paraxial_bwd_one(p, b, gout, gc, gt, gn);
}
#endif
