/*
 * torchoptics_b200 -- C ABI of the B200-native sequential ray-trace hot path.
 *
 * This is the drop-in boundary for the hot path of the reference tracer
 * (/root/reference/torchlens/ray_tracing_lite.py, "rtl" below).  Every entry
 * point takes plain device pointers, element strides and sizes; nothing here
 * knows about torch.  The caller owns all memory (inputs, outputs, workspace)
 * and passes the CUDA stream to launch on (a cudaStream_t cast to void*; NULL =
 * the legacy default stream).  All functions return 0 on success or a negative
 * TL_ERR_* code and never throw; tl_last_error() gives the message of the last
 * failure on the calling thread.  Nothing allocates, synchronises or copies to
 * the host (exception: tl_peer_create / tl_peer_destroy own the peer-exchange window).
 *
 * Tensor convention (rtl:4-9): ray tensors are [B lenses, F fields, P pupil
 * points, W wavelengths]; prescriptions carry a trailing surface axis S.
 *
 *   tl_trace_fwd           replaces  trace_skew forward                 rtl:594-675
 *                          (find_marching_distance_spherical rtl:525-545,
 *                           update_ray_coordinates rtl:514-522,
 *                           apply_snell_spherical rtl:548-571,
 *                           reset_bad_rays rtl:574-591)
 *   tl_trace_bwd           replaces  autograd of trace_skew             rtl:594-675
 *   tl_rms_fwd/tl_rms_bwd  replace   compute_rms2d and its autograd     rtl:678-702
 *   tl_spot_accumulate +   replace   trace_skew -> compute_rms2d -> .backward()
 *   tl_spot_finalize                 as one pass (the fwd+bwd hot path) rtl:594-702
 */
#ifndef TORCHOPTICS_B200_H
#define TORCHOPTICS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TL_ABI_VERSION 12

enum {
  TL_OK = 0,
  TL_ERR_INVALID = -1,      /* bad argument (null pointer, size, unsupported S) */
  TL_ERR_WORKSPACE = -2,    /* workspace too small                             */
  TL_ERR_CUDA = -3          /* a CUDA runtime call / launch failed             */
};

/* Arithmetic policy of the forward trace.
 * TL_ARITH_GUARDED: FMA-contracted fast path for rays that stay clear of every
 *   mask threshold by a guard band; every other ray is re-traced with
 *   TL_ARITH_EXACT arithmetic, so masks are those of the exact path.
 * TL_ARITH_EXACT: every operation individually rounded (IEEE-754 RN, no
 *   contraction) in the operand order of rtl:525-571 -- outputs are bit-identical
 *   to the reference's eager fp32 evaluation. */
enum { TL_ARITH_GUARDED = 0, TL_ARITH_EXACT = 1 };

#define TL_MAX_SURFACES_FWD   256  /* tl_trace_fwd                          */
#define TL_MAX_SURFACES_BWD    32  /* tl_trace_bwd                          */
#define TL_MAX_SURFACES_SPOT   32  /* tl_spot_accumulate with want_grad (17..32: 4-8 warps per SM) */
#define TL_MAX_SURFACES_GEN    48  /* any entry point, general surfaces     */

/* A ray-bundle input broadcastable to [B,F,P,W]: base pointer + element strides
 * (0 = broadcast along that axis), exactly what torch.broadcast_to() yields. */
typedef struct TlStrided {
  const float *ptr;
  int64_t stride[4];
} TlStrided;

/* Arguments of trace_skew (rtl:594). */
typedef struct TlProblem {
  TlStrided x, y, z, cx, cy;   /* entrance-pupil point and direction cosines      */
  const float *c;              /* [B,S] surface curvatures                         */
  const float *t;              /* [B,S] thickness after each surface               */
  const float *mu;             /* [B,W,S] index ratio n/n' per wavelength          */
  const uint8_t *live;         /* [B,S] structure mask (1 = real surface)          */
  int32_t B, F, P, W, S;
  int32_t allow_backward_rays; /* rtl:594 flag                                     */
  int32_t arith;               /* TL_ARITH_*                                       */
  int32_t p_begin, p_end;      /* pupil slice [p_begin,p_end) traced by this call
                                  (tl_spot_accumulate only; shards rays over GPUs) */
  const float *xy_scale;       /* [B] or NULL: x and y are multiplied by xy_scale[b] on load
                                  (relative pupil coordinates * EPD/2, scale_to_epd rtl:497-507) */
  /* EXTENSION surfaces (no reference behaviour; defined by oracle/asphere_oracle.py).  All three
   * NULL = the reference's spherical lens.  Any non-NULL selects the general-surface kernels:
   *   z = c rho / (1 + sqrt(1 - (1+k) c^2 rho)) + sum_i a_i rho^i,  rho = x^2 + y^2, i = 2..8
   * The intersection is the oracle's: base-sphere closed form, then Newton on F(tau) = z + tau cz - s(rho) --
   * four fixed steps under TL_ARITH_EXACT (bit-identical to the oracle); under TL_ARITH_GUARDED the loop is
   * left once a step is <= 1e-3 |tau| (the remaining ones would move tau by less than fp32 resolution). */
  const float *k;              /* [B,S]   conic constants (NULL = 0)                */
  const float *a;              /* [B,S,7] a4, a6, ..., a16 (NULL = 0)               */
  const float *sd;             /* [B,S]   clear semi-diameters (NULL = +inf)        */
  /* Ray-aiming map of tl_aim, [B,F,W,3] = (x_gain, y_gain, y_shift), or NULL: x and y are then
   * RELATIVE pupil coordinates and every kernel forms clamp(x * x_gain, -2, 2) * xy_scale[b] and
   * clamp(y * y_gain + y_shift, -2, 2) * xy_scale[b] on load (rtl:109-114), so that an aimed ray
   * set never has to be materialised as [B,F,P,W] tensors.  Spherical lenses only; not together
   * with per-ray gradients of x / y. */
  const float *aim;
  /* Pupil vignetting (apply_vignetting, ray_tracing.py:479-490 of the TensorFlow original; use sites
   * rtl:98-104, :154-160), [B,F,3] = (x_scale, y_scale, y_offset), or NULL: the relative pupil
   * coordinates become x * x_scale and y * y_scale + y_offset BEFORE the ray-aiming map, with
   * x_scale = 1 - vig_x, y_scale = 1 - (vig_up + vig_down) / 2, y_offset = (vig_down - vig_up) / 2
   * evaluated by the caller's vignetting function per (lens, field).  Same kernels as `aim`. */
  const float *vig;
} TlProblem;

/* Outputs of trace_skew, each a contiguous [B,F,P,W] array. */
typedef struct TlTraceOut {
  float *x, *y, *cx, *cy;
  uint8_t *ok, *backward;      /* 0/1 bytes (torch.bool storage)                   */
  float *opl;                  /* optional: optical path length entrance -> image
                                  (general-surface lenses only; NULL = not wanted);
                                  differentiable: TlSeeds.gopl                      */
  /* trace_skew(aggregate=True), rtl:641-657: the three penalty stacks, each a contiguous
   * [S,B,F,P,W] array (surface-major; the reference returns them as S-long lists).  All three
   * NULL = aggregate=False; spherical lenses only.
   *   z_relu      = max(z, 0) of the ray behind surface k, in the next vertex frame
   *   theta       = acos(clamp(cos theta))  / (pi/2), 1 for a ray that is not ok behind surface k
   *   theta_prime = acos(clamp(cos theta')) / (pi/2), likewise                                   */
  float *z_relu, *theta, *theta_prime;
} TlTraceOut;

/* Upstream gradients of the four differentiable outputs; contiguous [B,F,P,W]
 * or NULL (= zero). */
typedef struct TlSeeds {
  const float *gx, *gy, *gcx, *gcy;
  /* upstream gradients of the aggregate=True stacks, contiguous [S,B,F,P,W] or NULL (= zero).
   * Where the reference's own gradient is NaN (it takes sqrt of a failed ray's negative cos^2
   * before masking it, rtl:646-654) a failed ray contributes 0 here. */
  const float *gz_relu, *gtheta, *gtheta_prime;
  /* upstream gradient of TlTraceOut.opl (general-surface lenses only), contiguous [B,F,P,W] or NULL.
   * A ray that is not ok contributes nothing (its opl is a partial sum up to the failure). */
  const float *gopl;
} TlSeeds;

/* Gradients produced by tl_trace_bwd.  Prescription gradients are summed over
 * rays; per-ray gradients are contiguous [B,F,P,W] and optional (NULL = skip). */
typedef struct TlGrads {
  float *gc;                   /* [B,S]                                            */
  float *gt;                   /* [B,S]                                            */
  float *gmu;                  /* [B,W,S]                                          */
  float *gz_sum;               /* [B]   d/dz summed over the lens' rays            */
  float *gx, *gy, *gz, *gcx, *gcy;  /* per ray, optional                           */
  float *gk, *ga;              /* [B,S], [B,S,7]: general-surface lenses (required there)  */
} TlGrads;

/* Results of the fused spot pass for each lens. */
typedef struct TlSpotOut {
  float *rms;                  /* [B]   mean over fields of the y-RMS (rtl:678-702, every lens) */
  float *rms_field;            /* [B,F] per-field RMS                               */
  float *gc, *gt, *gmu, *gz;   /* d rms[b] / d{c[b,:], t[b,:], mu[b,:,:], z[b]}; NULL if no grad */
  float *gk, *ga;              /* general-surface lenses: d rms[b] / d{k[b,:], a[b,:,:]}; both
                                  non-NULL selects the general moment layout          */
} TlSpotOut;

int tl_abi_version(void);
const char *tl_last_error(void);

/* Forward trace: fills all six outputs. */
int tl_trace_fwd(const TlProblem *pb, const TlTraceOut *out, void *stream);

/* Backward of the trace.  Re-traces every ray (nothing is saved by the forward),
 * applies the adjoint surface by surface and reduces the prescription gradients
 * over rays in fp64.  Workspace: tl_trace_bwd_workspace() bytes, 8-byte aligned. */
size_t tl_trace_bwd_workspace(const TlProblem *pb);
int tl_trace_bwd(const TlProblem *pb, const TlSeeds *seeds, const TlGrads *grads,
                 void *workspace, size_t workspace_bytes, void *stream);

/* compute_rms2d for every lens of the batch and its backward.
 * y, ok: contiguous [B,F,P,W].  stats: [B,F,4] doubles written by the forward
 * and consumed by the backward.  gy = grad_rms[b] * d rms[b] / d y. */
size_t tl_rms_workspace(int32_t B, int32_t F, int32_t P, int32_t W);
int tl_rms_fwd(const float *y, const uint8_t *ok, int32_t B, int32_t F, int32_t P, int32_t W,
               float *rms, float *rms_field, double *stats,
               void *workspace, size_t workspace_bytes, void *stream);
int tl_rms_bwd(const float *y, const uint8_t *ok, const double *stats, const float *grad_rms,
               int32_t B, int32_t F, int32_t P, int32_t W, float *gy, void *stream);

/* Fused hot path.  tl_spot_accumulate traces the pupil slice [p_begin,p_end) of
 * every (lens, field, wavelength), and in the same pass applies the adjoint with
 * a unit seed on y so that the RMS gradient can be assembled later without a
 * second trace.  It leaves, in `moments` ([B,F,W,tl_spot_moment_count()] doubles),
 * sums that are ADDITIVE over pupil slices: ranks that traced different slices
 * all-reduce (sum) this one buffer.  `ref_y` ([B,F] floats) receives the per-field
 * reference height the sums are centred on (identical on every rank).
 * tl_spot_finalize turns the (reduced) moments into the RMS and its gradients;
 * P_total is the full pupil count of the job. */
int32_t tl_spot_moment_count(int32_t S, int32_t want_grad);
int32_t tl_spot_moment_count_general(int32_t S, int32_t want_grad);   /* general-surface lenses */
size_t tl_spot_workspace(const TlProblem *pb, int32_t want_grad);
int tl_spot_accumulate(const TlProblem *pb, int32_t want_grad, double *moments, float *ref_y,
                       void *workspace, size_t workspace_bytes, void *stream);
int tl_spot_finalize(const double *moments, const float *ref_y, int32_t B, int32_t F, int32_t W,
                     int32_t S, int64_t P_total, int32_t want_grad, const TlSpotOut *out,
                     void *stream);
/* Name of the kernel tl_spot_accumulate would launch for this problem (diagnostics / bench.py;
 * thread-local storage, valid until the next call). */
const char *tl_spot_kernel_name(const TlProblem *pb, int32_t want_grad);
/* The dominant kernel of tl_spot_accumulate (want_grad) ALONE -- no reference-height launch, no row
 * reduction; its partial sums stay in the workspace -- so that bench.py can time exactly one kernel
 * with CUDA events.  Spherical problems that take k_spot_rev only. */
int tl_spot_kernel_only(const TlProblem *pb, const float *ref_y, void *workspace, size_t workspace_bytes,
                        void *stream);

/* Fused penalty pass: value and gradient of the ray-angle / ray-path penalty that compute_loss_out
 * (optics_simulator_lite.py:430-450) builds from trace_skew(aggregate=True):
 *     penalty[b] = scale * sum over rays and surfaces of (theta_norm + theta_prime_norm + z_RELU),
 * scale = 1 / numSequence, as ONE pass that materialises no stack (the reference stacks 3 S
 * [B,F,P,W] tensors).  Like the spot pass it is split so that ranks can shard the pupil:
 * tl_penalty_accumulate traces the slice [p_begin,p_end) and leaves sums that are additive over
 * slices in `moments` ([B,F,W,tl_penalty_moment_count(S)] doubles; all-reduce them);
 * tl_penalty_finalize turns them into the penalty and d penalty / d{c, t, mu, z}.  Failed rays
 * count as the reference counts them (theta = theta' = 1, z = -t) and contribute no gradient
 * (the reference's gradient is NaN there, see TlSeeds). */
typedef struct TlPenaltyOut {
  float *penalty;              /* [B]                                                */
  float *gc, *gt, *gmu, *gz;   /* [B,S], [B,S], [B,W,S], [B]                          */
} TlPenaltyOut;
int32_t tl_penalty_moment_count(int32_t S);
size_t tl_penalty_workspace(const TlProblem *pb);
int tl_penalty_accumulate(const TlProblem *pb, double *moments, void *workspace, size_t workspace_bytes,
                          void *stream);
int tl_penalty_finalize(const double *moments, int32_t B, int32_t F, int32_t W, int32_t S, double scale,
                        const TlPenaltyOut *out, void *stream);

/* Ray-set staging of RayTracer.trace_rays (rtl:80-124) and its chain rule, as two small
 * kernels instead of ~60 eager tensor ops: the two-term dispersion model n(lambda) of
 * Lens.get_refractive_indices (lens_modeling.py:355-374), the index ratios mu = n/n'
 * (rtl:123), the paraxial entrance-pupil position z = B/A of the ABCD product in front of
 * the stop (compute_pupil_position rtl:330-350) and cy = sin(hfov * rel_field) (rtl:116).
 * Prescriptions are the padded [B,L] tensors of lens_modeling.Lens. */
typedef struct TlLens {
  const float *c, *t, *nd, *v;        /* [B,L]                                       */
  const uint8_t *mask, *mask_g;       /* [B,L] surface present / followed by glass   */
  const int32_t *stop_idx;            /* [B]                                         */
  const float *hfov;                  /* [B] half field of view [rad]                */
  const float *epd;                   /* [B] entrance pupil diameter                 */
  const float *rel_fields;            /* [F]                                         */
  const float *wavelengths;           /* [W] nm                                      */
  int32_t B, L, F, W;
} TlLens;

/* mu [B,W,L], z [B], cy [B,F], half_epd [B] (= the xy_scale of TlProblem) */
int tl_stage_fwd(const TlLens *lens, float *mu, float *z, float *cy, float *half_epd, void *stream);
/* tl_stage_fwd, (aim != NULL) tl_aim and (ref_y != NULL) the reference heights of tl_spot_accumulate as
 * ONE launch: `pb` is the problem the fused pass will run (its mu, z, cy, xy_scale, aim pointers
 * are the very buffers written here; read only when ref_y != NULL).  Pair it with
 * tl_spot_accumulate_ref, which takes ref_y as an input. */
int tl_stage_ref(const TlLens *lens, const TlProblem *pb, float *mu, float *z, float *cy, float *half_epd,
                 float *aim, const float *vig, int32_t aim_mode, int32_t allow_backward_rays, float *ref_y,
                 void *stream);
int tl_spot_accumulate_ref(const TlProblem *pb, int32_t want_grad, double *moments, const float *ref_y,
                           void *workspace, size_t workspace_bytes, void *stream);
/* tl_spot_finalize (with gradients) + the chain rule of tl_stage_bwd as ONE launch: from the (reduced)
 * moments to rms, rms_field and d rms[b] / d{c, t, nd, v} [B,L] (out->gc, out->gt, gnd, gv; written,
 * not added to).  out->gmu [B,W,L] and out->gz [B] are scratch. */
int tl_lens_spot_finalize(const double *moments, const float *ref_y, const TlLens *lens, int64_t P_total,
                          const TlSpotOut *out, float *gnd, float *gv, void *stream);

/* Chain rule: given d loss / d mu [B,W,L] and d loss / d z [B], ADDS the induced gradients to
 * gc, gt, gnd, gv [B,L] (which the caller pre-fills, e.g. with the direct c / t gradients). */
int tl_stage_bwd(const TlLens *lens, const float *gmu, const float *gz, float *gc, float *gt,
                 float *gnd, float *gv, void *stream);

/* Peer-memory exchange: the one collective of the pupil-sharded spot pass (SUM all-reduce of the
 * `moments` buffer of tl_spot_accumulate over the ranks of ONE node) as a single kernel that
 * stores into every peer's window over NVLink / NVSwitch and sums in rank order (bit-identical
 * result on every rank), instead of an NCCL call.  The reference is single-GPU (no counterpart).
 *   tl_peer_create   allocates this rank's window (the one allocation this library makes: CUDA IPC
 *                    needs a cudaMalloc base pointer) on the CURRENT device and writes its IPC
 *                    handle (tl_peer_handle_bytes() bytes) to handle_out.
 *   tl_peer_connect  takes the handles of all ranks (world x handle bytes, rank order; the caller
 *                    exchanges them, e.g. with torch.distributed.all_gather) and maps the windows.
 *   tl_peer_allreduce_f64  out[i] = sum over ranks of data[i], i < n <= capacity; every rank must
 *                    call it the same number of times; graph-capturable (no per-call state on the
 *                    host); data != out when world > 1.
 *                    Every all-reduce of one communicator must be issued on ONE stream (the epoch is
 *                    device state).  A wait for a peer that times out (4 s) is loud and sticky: this
 *                    and every later call write NaN to `out` and set the status word.
 *   tl_peer_status   (synchronising) status 0 = ok, 1 = a wait for a peer timed out. */
typedef struct TlPeerComm TlPeerComm;
size_t tl_peer_handle_bytes(void);
int tl_peer_create(int32_t rank, int32_t world, int64_t capacity, TlPeerComm **comm, void *handle_out);
int tl_peer_connect(TlPeerComm *comm, const void *all_handles);
int tl_peer_allreduce_f64(TlPeerComm *comm, const double *data, double *out, int64_t n, void *stream);
int tl_peer_status(TlPeerComm *comm, int32_t *status_out, uint32_t *epoch_out);
int tl_peer_destroy(TlPeerComm *comm);

/* Ray aiming (RayTracer.ray_aiming rtl:129-208; one iteration, ray_aiming_mode 'real', no pupil
 * vignetting function) as one kernel instead of three nested eager traces and two autograd
 * backward calls: from the staged mu, z, cy, half_epd of tl_stage_fwd it traces the marginal ray
 * (stop radius, compute_pupil_radius rtl:834-844) and the three 'tee' rays of every (lens, field,
 * wavelength) to the stop in forward mode and writes the affine pupil map
 *     aim[b,f,w] = (x_gain, y_gain, y_shift):  x_rel -> x_rel * x_gain,  y_rel -> y_rel * y_gain + y_shift
 * (rtl:196-206).  `vig` ([B,F,3] as TlProblem.vig, or NULL): the tee rays and their targets are the
 * vignetted ones (rtl:154-160).  `aim_mode`: TL_AIM_REAL takes the stop radius from the traced
 * marginal ray, TL_AIM_PARAXIAL from the first-order magnification of the front group times EPD / 2
 * (compute_magnification, ray_tracing.py:765-777; rtl:138-140).  Like the reference's (it traces a
 * detached lens, rtl:111) the map carries no gradient. */
enum { TL_AIM_REAL = 0, TL_AIM_PARAXIAL = 1 };
int tl_aim(const TlLens *lens, const float *mu, const float *z, const float *cy, const float *half_epd,
           const float *vig, int32_t aim_mode, int32_t allow_backward_rays, float *aim, void *stream);

/* Spot-diagram / PSF binning: the Gaussian soft histogram of the reference's compute_psf
 * (ray_tracing.py:206-270, the TensorFlow original).  Rays are given per
 * grid g (= lens x field) and colour channel c; bin centres and sigma = half a bin as rt_tf:238-247;
 * only the non-negative half of the x bins is evaluated (n_xh = n_x_bins / 2, + 1 if odd): the caller
 * mirrors and normalises (rt_tf:257-263).
 *   sums      [G, C, n_y_bins, n_xh] doubles: the un-normalised histogram
 *   inside    [G, C] doubles: rays with |y - y_target| < y_size / 2 and |x| < x_size / 2 (rt_tf:266-267)
 * Workspace: tl_psf_workspace() bytes. */
typedef struct TlPsf {
  const float *x, *y;              /* [G, C, R] image-plane points                            */
  const float *y_target;           /* [G] grid centre in y (x is centred on 0)                */
  const float *x_incr, *y_incr;    /* [G] bin pitch                                           */
  const float *x_size, *y_size;    /* [G] window extent of the accounted-ray count            */
  int32_t G, C, R;
  int32_t n_x_bins, n_y_bins;
} TlPsf;
size_t tl_psf_workspace(const TlPsf *psf);
int tl_psf_bin(const TlPsf *psf, double *sums, double *inside, void *workspace, size_t workspace_bytes,
               void *stream);

/* Paraxial (ABCD) front end of a padded lens batch, one thread per lens:
 * `get_first_order` rtl:772-794 and `compute_last_curvature` rtl:725-769 (over
 * `interface_propagation_abcd` rtl:314-327 / `reduce_abcd` rtl:301-311), which the reference's
 * consumer runs once per sample in a Python loop (optical_loss.py:63-64, :99-115).
 *   c, t [B,L]; n [B,L] = refractive index BEHIND each slot (1 behind air slots and padding);
 *   live, glass [B,L] bytes = Structure.mask / mask_G (prefix masks, as the reference assumes).
 * tl_paraxial_fwd writes out [B,2]:
 *   TL_PARAXIAL_FIRST_ORDER      (EFL, BFL), the last live slot's thickness taken as 0 (rtl:781-783)
 *   TL_PARAXIAL_LAST_CURVATURE   (the curvature that makes EFL = 1, the slot it belongs to): the slot is the
 *                                last one, or the one before it when the sequence ends air-air
 *                                (rtl:735-738); c at and behind that slot is not read.
 * tl_paraxial_bwd: gout [B,2] (column 1 of LAST_CURVATURE ignored) -> gc, gt, gn [B,L], every element
 * written.  L <= TL_PARAXIAL_MAX_SLOTS. */
#define TL_PARAXIAL_MAX_SLOTS 64
enum { TL_PARAXIAL_FIRST_ORDER = 0, TL_PARAXIAL_LAST_CURVATURE = 1 };
typedef struct TlParaxial {
  const float *c, *t, *n;          /* [B,L] */
  const uint8_t *live, *glass;     /* [B,L] */
  int32_t B, L;
  int32_t mode;
} TlParaxial;
int tl_paraxial_fwd(const TlParaxial *lens, float *out, void *stream);
int tl_paraxial_bwd(const TlParaxial *lens, const float *gout, float *gc, float *gt, float *gn, void *stream);

/* Layout self-description, so that a binding can check itself against the library it loaded:
 * for struct `which` (0 TlStrided, 1 TlProblem, 2 TlTraceOut, 3 TlSeeds, 4 TlGrads, 5 TlSpotOut,
 * 6 TlPenaltyOut, 7 TlLens, 8 TlPsf, 9 TlParaxial) returns "Name:sizeof;field@offsetof;field@offsetof;..." in declaration
 * order (thread-local storage, valid until the next call), or NULL for an unknown struct. */
const char *tl_abi_describe(int32_t which);

/* Number of kernels this library has launched since it was loaded (bench.py's
 * gpu_launches counter). */
int64_t tl_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TORCHOPTICS_B200_H */
