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
// input: for the reference's 21 x 21 grid that is 32 MUFU + 231 FMA per 8 B (590 flop per ray counting 4 per
// exponential) -- compute-bound on the FP32 / MUFU pipes by a wide margin (the HBM roofline would be
// 0.8 T rays/s, the FP32 one 126 G rays/s, the MUFU one 145 G rays/s), not a dense
// contraction worth reshaping for the tensor cores (K = rays, but M x N = 21 x 11 and both operands
// would have to be materialised first: 128 B of exponentials per ray).
//
// A CTA takes one (grid, channel) and a contiguous chunk of its rays, in tiles of kPsfTile rays:
//   phase 1  the separable factors of the tile, computed once into shared memory (threads over rays x
//            bin rows; rows padded to a multiple of the register block with zeros);
//   phase 2  REGISTER-BLOCKED accumulation: a thread owns a 3 x 4 block of bins and a subset of the
//            tile's rays, four consecutive rays per step: 3 + 4 128-bit shared loads feed 48 FMAs
//            (the first version gave a thread one bin: 2 scalar shared loads per FMA, shared-memory
//            bound at 7 % of the FP32 peak);
//   end      the ray subsets of a bin block are summed through shared memory; chunk sums leave as fp64
//            partials and k_psf_reduce adds them in a fixed order (deterministic).

constexpr int kPsfThreads = 256;
constexpr int kPsfTile = 256;            // rays per tile = threads per CTA (phase 1: a ray per thread)
constexpr int kPsfStride = kPsfTile + 4; // floats per factor row: 16-byte aligned, rows shifted by 4 banks
constexpr int kPsfMaxBins = 64;          // per axis
constexpr int kPsfBy = 3, kPsfBx = 4;    // register block of bins per thread
static_assert(kPsfTile == kPsfThreads, "phase 1 gives every thread one ray of the tile");

struct PsfArgs {
  TlPsf p;
  double *partial;        // [G * C, n_chunks, n_y * n_xh + 1]   (+1: rays inside the window)
  int n_chunks, chunk_len, n_xh;
};

__global__ void __launch_bounds__(kPsfThreads)
k_psf_bin(PsfArgs a) {
  extern __shared__ __align__(16) float sm[];
  const TlPsf &p = a.p;
  const int n_xh = a.n_xh, n_y = p.n_y_bins, n_bins = n_xh * n_y;
  const int n_by = (n_y + kPsfBy - 1) / kPsfBy, n_bx = (n_xh + kPsfBx - 1) / kPsfBx;
  const int rows_y = n_by * kPsfBy, rows_x = n_bx * kPsfBx;          // padded row counts
  const int n_blocks = n_by * n_bx;                                    // bin blocks (<= 22 * 8 = 176)
  const int n_sub = kPsfThreads / n_blocks;                            // ray subsets per bin block
  float *gy = sm;                                    // [rows_y][kPsfStride]
  float *gx = sm + (size_t)rows_y * kPsfStride;      // [rows_x][kPsfStride]
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
  __shared__ float centres[2 * kPsfMaxBins + kPsfBy + kPsfBx];      // bin-row centres: y rows, then x rows
  if (threadIdx.x == 0) inside_total = 0.f;
  for (int row = threadIdx.x; row < rows_y + rows_x; row += kPsfThreads)
    centres[row] = row < rows_y ? ((float)row + y_first) * y_incr : ((float)(row - rows_y) + x_first) * x_incr;
  float inside = 0.f;
  const int block_id = threadIdx.x % n_blocks, sub = threadIdx.x / n_blocks;
  const bool worker = sub < n_sub;
  const int by = block_id / n_bx, bx = block_id % n_bx;
  const float *fy = gy + (size_t)by * kPsfBy * kPsfStride, *fx = gx + (size_t)bx * kPsfBx * kPsfStride;
  float acc[kPsfBy][kPsfBx];
#pragma unroll
  for (int i = 0; i < kPsfBy; ++i)
#pragma unroll
    for (int j = 0; j < kPsfBx; ++j) acc[i][j] = 0.f;

  for (int r0 = r_lo; r0 < r_hi; r0 += kPsfTile) {
    const int n = min(kPsfTile, r_hi - r0);
    __syncthreads();                                  // the previous tile's readers are done
    // phase 1: a thread takes ONE ray of the tile (two coalesced loads) and walks the bin rows: per factor a
    // broadcast load of the row centre, subtract, scale, square, ex2, store.  Rows past the bin count and rays
    // past the tile hold zeros.  (ex2.approx.ftz: arguments <= 0; results below 2^-126 flush to 0, which is what
    // they contribute to a sum of O(1) terms.)
    {
      const int t = threadIdx.x;            // kPsfTile == kPsfThreads
      const bool has = t < n;
      const float xr = has ? xs[r0 + t] : 0.f;
      const float yr = has ? ys[r0 + t] - yt : 0.f;
      inside += (has && fabsf(yr) < half_y && fabsf(xr) < half_x) ? 1.f : 0.f;      // rays inside the window (rt_tf:266-267)
      for (int row = 0; row < rows_y; ++row) {
        const float d = (yr - centres[row]) * ky;
        float v;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(-0.72134752044448170f * d * d));
        gy[(size_t)row * kPsfStride + t] = (has && row < n_y) ? v : 0.f;
      }
      for (int row = 0; row < rows_x; ++row) {
        const float d = (xr - centres[rows_y + row]) * kx;
        float v;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(-0.72134752044448170f * d * d));
        gx[(size_t)row * kPsfStride + t] = (has && row < n_xh) ? v : 0.f;
      }
    }
    __syncthreads();
    // phase 2: this thread's ray quads of the tile
    if (worker) {
      for (int q = sub; q < kPsfTile / 4; q += n_sub) {
        float4 vy[kPsfBy], vx[kPsfBx];
#pragma unroll
        for (int i = 0; i < kPsfBy; ++i) vy[i] = *reinterpret_cast<const float4 *>(fy + (size_t)i * kPsfStride + 4 * q);
#pragma unroll
        for (int j = 0; j < kPsfBx; ++j) vx[j] = *reinterpret_cast<const float4 *>(fx + (size_t)j * kPsfStride + 4 * q);
#pragma unroll
        for (int i = 0; i < kPsfBy; ++i)
#pragma unroll
          for (int j = 0; j < kPsfBx; ++j) {
            float s = acc[i][j];
            s = fmaf(vy[i].x, vx[j].x, s);
            s = fmaf(vy[i].y, vx[j].y, s);
            s = fmaf(vy[i].z, vx[j].z, s);
            s = fmaf(vy[i].w, vx[j].w, s);
            acc[i][j] = s;
          }
      }
    }
  }
  // sum the ray subsets of every bin block: scratch[sub][block][3][4] over the factor rows
  __syncthreads();
  float *scratch = sm;
  if (worker) {
#pragma unroll
    for (int i = 0; i < kPsfBy; ++i)
#pragma unroll
      for (int j = 0; j < kPsfBx; ++j)
        scratch[((size_t)sub * n_blocks + block_id) * (kPsfBy * kPsfBx) + i * kPsfBx + j] = acc[i][j];
  }
  __syncthreads();
  double *dst = a.partial + ((int64_t)gc * a.n_chunks + chunk) * (n_bins + 1);
  for (int bin = threadIdx.x; bin < n_bins; bin += kPsfThreads) {
    const int iy = bin / n_xh, ix = bin % n_xh;
    const int blk = (iy / kPsfBy) * n_bx + ix / kPsfBx, slot = (iy % kPsfBy) * kPsfBx + ix % kPsfBx;
    double s = 0.0;
    for (int u = 0; u < n_sub; ++u) s += (double)scratch[((size_t)u * n_blocks + blk) * (kPsfBy * kPsfBx) + slot];
    dst[bin] = s;
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
  const size_t rows = (size_t)((p.n_y_bins + kPsfBy - 1) / kPsfBy) * kPsfBy + (size_t)((pl.n_xh + kPsfBx - 1) / kPsfBx) * kPsfBx;
  pl.smem = rows * kPsfStride * sizeof(float);
  const size_t scratch = (size_t)kPsfThreads * kPsfBy * kPsfBx * sizeof(float);      // the end-of-chunk sum reuses the rows
  if (pl.smem < scratch) pl.smem = scratch;
  const int64_t n_gc = (int64_t)p.G * p.C;
  const int64_t tiles = ((int64_t)p.R + kPsfTile - 1) / kPsfTile;
  // ONE wave: as many chunks per (grid, channel) as fit the resident CTAs of the device (a 5 % overshoot of the
  // residency -- 624 CTAs on 592 slots -- doubled the run time of the first version)
  if (pl.smem > 48 * 1024)
    TL_CHECK_CUDA(cudaFuncSetAttribute((const void *)k_psf_bin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  int per_sm = 1;
  TL_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)k_psf_bin, kPsfThreads, pl.smem));
  if (per_sm < 1) per_sm = 1;
  int64_t chunks = (int64_t)info.sms * per_sm / n_gc;
  if (chunks > tiles) chunks = tiles;
  if (chunks < 1) chunks = 1;
  const int64_t tiles_per_chunk = (tiles + chunks - 1) / chunks;
  pl.chunk_len = (int)(tiles_per_chunk * kPsfTile);
  pl.n_chunks = (int)(((int64_t)p.R + pl.chunk_len - 1) / pl.chunk_len);
  pl.partial_bytes = align8((size_t)n_gc * pl.n_chunks * ((size_t)pl.n_xh * p.n_y_bins + 1) * sizeof(double));
  return TL_OK;
}
