// Spot-diagram / PSF binning: the Gaussian soft histogram of the reference's compute_psf
// (/root/reference/torchlens/ray_tracing.py:206-270 -- the TensorFlow original; the function is absent
// from ray_tracing_lite.py and its consumer sample_psfs is commented out at
// optics_simulator_lite.py:656-677).  SURVEY.md section 8(f)-4.  Included by trace_kernels.cu.
//
//   kernels[g, c, iy, ix] = sum over rays r of  exp(-(x_r - X_ix)^2 / 2 sx^2) * exp(-(y_r - yt_g - Y_iy)^2 / 2 sy^2)
//
// with sigma = half a bin (rt_tf:246-247), only the non-negative half of the x bins evaluated
// (rt_tf:238-243; the caller mirrors and normalises, rt_tf:257-263).  Every ray touches every bin,
// so per ray the work is (n_xh + n_y) exponentials and n_xh * n_y multiply-adds against 8 bytes of
// input: for the reference's 21 x 21 grid that is 32 MUFU + 231 FMA per 8 B -- compute-bound on the
// FP32 / MUFU pipes by a wide margin (the HBM roofline would be 0.8 T rays/s), not a dense
// contraction worth reshaping for the tensor cores (K = rays, but M x N = 21 x 11 and both operands
// would have to be materialised first: 128 B of exponentials per ray).
//
// A CTA takes one (grid, channel) and a contiguous chunk of its rays: tiles of kPsfTile rays, the
// separable factors of a tile computed once into shared memory (threads over rays x bins), then one
// accumulator per (iy, ix) bin and thread, the ray index running over the tile -- the factor rows are
// read as broadcasts (same ix or iy across a quarter warp) from rows padded to an odd stride.  Chunk
// sums leave as fp64 partials; k_psf_reduce adds them in a fixed order (deterministic).

constexpr int kPsfThreads = 256;
constexpr int kPsfTile = 128;
constexpr int kPsfMaxBins = 64;          // per axis

struct PsfArgs {
  TlPsf p;
  double *partial;        // [G * C, n_chunks, n_y * n_xh + 1]   (+1: rays inside the window)
  int n_chunks, chunk_len, n_xh;
};

__global__ void __launch_bounds__(kPsfThreads)
k_psf_bin(PsfArgs a) {
  extern __shared__ float sm[];
  const TlPsf &p = a.p;
  const int n_xh = a.n_xh, n_y = p.n_y_bins, n_bins = n_xh * n_y;
  constexpr int kStride = kPsfTile + 1;
  float *gx = sm;                        // [n_xh][kStride]
  float *gy = sm + (size_t)n_xh * kStride;      // [n_y][kStride]
  __shared__ float inside_total;
  const int chunk = blockIdx.x % a.n_chunks;
  const int gc = blockIdx.x / a.n_chunks;
  const int g = gc / p.C;
  const float yt = p.y_target[g];
  const float x_incr = p.x_incr[g], y_incr = p.y_incr[g];
  // exp(-(d / sigma)^2 / 2) = exp2(-(log2 e / 2) (d / sigma)^2), sigma = incr / 2
  const float kx = 2.0f / x_incr, ky = 2.0f / y_incr;
  const float x_first = (p.n_x_bins % 2 == 1) ? 0.0f : 0.5f;                       // rt_tf:239-242
  const float y_first = 0.5f - 0.5f * (float)n_y;                                  // rt_tf:243
  const float half_x = 0.5f * p.x_size[g], half_y = 0.5f * p.y_size[g];
  const float *xs = p.x + (int64_t)gc * p.R, *ys = p.y + (int64_t)gc * p.R;
  const int r_lo = chunk * a.chunk_len, r_hi = min(p.R, r_lo + a.chunk_len);
  if (threadIdx.x == 0) inside_total = 0.f;
  float inside = 0.f;
  constexpr int kPerThread = (kPsfMaxBins * kPsfMaxBins / 2 + kPsfThreads - 1) / kPsfThreads;    // <= 8 bins / thread
  float acc[kPerThread];
#pragma unroll
  for (int q = 0; q < kPerThread; ++q) acc[q] = 0.f;
  for (int r0 = r_lo; r0 < r_hi; r0 += kPsfTile) {
    const int n = min(kPsfTile, r_hi - r0);
    __syncthreads();                                  // the previous tile's readers are done
    // separable factors of this tile: thread = (bin row, ray)
    for (int i = threadIdx.x; i < (n_xh + n_y) * kPsfTile; i += kPsfThreads) {
      const int row = i / kPsfTile, t = i % kPsfTile;
      float v = 0.f;
      if (t < n) {
        if (row < n_xh) {
          const float d = (xs[r0 + t] - ((float)row + x_first) * x_incr) * kx;
          v = exp2f(-0.72134752044448170f * d * d);
        } else {
          const float d = ((ys[r0 + t] - yt) - ((float)(row - n_xh) + y_first) * y_incr) * ky;
          v = exp2f(-0.72134752044448170f * d * d);
        }
      }
      (row < n_xh ? gx + (size_t)row * kStride : gy + (size_t)(row - n_xh) * kStride)[t] = v;
    }
    // rays inside the window (rt_tf:266-267)
    for (int t = threadIdx.x; t < n; t += kPsfThreads)
      inside += (fabsf(ys[r0 + t] - yt) < half_y && fabsf(xs[r0 + t]) < half_x) ? 1.f : 0.f;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int bin = threadIdx.x + q * kPsfThreads;
      if (bin >= n_bins) break;
      const float *fx = gx + (size_t)(bin % n_xh) * kStride, *fy = gy + (size_t)(bin / n_xh) * kStride;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
      for (int t = 0; t < kPsfTile; t += 4) {          // (the tail of a short tile holds zeros)
        s0 = fmaf(fx[t], fy[t], s0);
        s1 = fmaf(fx[t + 1], fy[t + 1], s1);
        s2 = fmaf(fx[t + 2], fy[t + 2], s2);
        s3 = fmaf(fx[t + 3], fy[t + 3], s3);
      }
      acc[q] += (s0 + s1) + (s2 + s3);
    }
  }
  double *dst = a.partial + ((int64_t)gc * a.n_chunks + chunk) * (n_bins + 1);
#pragma unroll
  for (int q = 0; q < kPerThread; ++q) {
    const int bin = threadIdx.x + q * kPsfThreads;
    if (bin < n_bins) dst[bin] = (double)acc[q];
  }
  inside = warp_sum(inside);
  if ((threadIdx.x & 31) == 0) atomicAdd(&inside_total, inside);      // (integers < 2^24: exact, any order)
  __syncthreads();
  if (threadIdx.x == 0) dst[n_bins] = (double)inside_total;
}

// partial[gc, chunk, n_bins + 1] -> sums[gc, n_bins], inside[gc] (fixed order)
__global__ void k_psf_reduce(const double *partial, double *sums, double *inside, int n_gc, int n_chunks, int n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_gc * n) return;
  const int gc = (int)(i / n), slot = (int)(i % n);
  double s0 = 0.0, s1 = 0.0;
  int c = 0;
  for (; c + 2 <= n_chunks; c += 2) {
    s0 += partial[((int64_t)gc * n_chunks + c) * n + slot];
    s1 += partial[((int64_t)gc * n_chunks + c + 1) * n + slot];
  }
  if (c < n_chunks) s0 += partial[((int64_t)gc * n_chunks + c) * n + slot];
  if (slot < n - 1) sums[(int64_t)gc * (n - 1) + slot] = s0 + s1;
  else inside[gc] = s0 + s1;
}

struct PsfPlan {
  int n_xh = 1, n_chunks = 1, chunk_len = 1;
  size_t smem = 0, partial_bytes = 0;
};

int plan_psf(const TlPsf &p, PsfPlan &pl) {
  DeviceInfo info;
  int rc = device_info(info);
  if (rc) return rc;
  if (p.G < 1 || p.C < 1 || p.R < 1 || p.n_x_bins < 1 || p.n_y_bins < 1 || p.n_x_bins > kPsfMaxBins ||
      p.n_y_bins > kPsfMaxBins)
    return fail(TL_ERR_INVALID, "tl_psf: G, C, R >= 1 and 1 <= bins per axis <= 64%s");
  pl.n_xh = p.n_x_bins % 2 == 1 ? p.n_x_bins / 2 + 1 : p.n_x_bins / 2;
  pl.smem = (size_t)(pl.n_xh + p.n_y_bins) * (kPsfTile + 1) * sizeof(float);
  const int64_t n_gc = (int64_t)p.G * p.C;
  const int64_t tiles = ((int64_t)p.R + kPsfTile - 1) / kPsfTile;
  int64_t chunks = ((int64_t)info.sms * 4 + n_gc - 1) / n_gc;        // ~4 CTAs per SM over the whole grid
  if (chunks > tiles) chunks = tiles;
  if (chunks < 1) chunks = 1;
  const int64_t tiles_per_chunk = (tiles + chunks - 1) / chunks;
  pl.chunk_len = (int)(tiles_per_chunk * kPsfTile);
  pl.n_chunks = (int)(((int64_t)p.R + pl.chunk_len - 1) / pl.chunk_len);
  pl.partial_bytes = align8((size_t)n_gc * pl.n_chunks * ((size_t)pl.n_xh * p.n_y_bins + 1) * sizeof(double));
  return TL_OK;
}
