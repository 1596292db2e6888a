// Covariance-matrix builders (HBM-write-bound kernels).
//
// Replaces (reference, moe/optimal_learning/cpp/):
//   BuildCovarianceMatrixWithNoiseVariance   gpp_math.cpp:426-455   (lower triangle + noise by observation type)
//   BuildMixCovarianceMatrix                 gpp_math.cpp:309-335
//   SquareExponential / MaternNu2p5 ::Covariance   gpp_covariance.cpp:121-164, 339-387
//
// Tiling: one CTA per 128x32 block of point pairs in the lower-triangular tile grid; the two point slabs are staged
// through shared memory once and every thread produces 4 consecutive rows of a column, so each warp store instruction
// writes 1 KB contiguous (coalesced 32 B per lane).  Only lower-triangular tiles are visited: the algorithmic traffic
// is 4*n*(n+1) bytes written + 8*N*dim read.
#include "device_math.cuh"
#include "internal.cuh"

namespace cmoe {

namespace {

constexpr int TR = 128;  // point rows per tile
constexpr int TC = 32;   // point cols per sub-tile
constexpr int TCW = 64;  // point cols per CTA tile (TCW / TC sub-tiles share the staged row slab)

// g == 0 fast path: K[i + j*n] for i >= j (tile granularity), noise[0] on the diagonal.
// Works on length-scaled coordinates Xs = X / l (so r^2 is a plain sum of squared differences: 2 FP64 ops per
// dimension, no division) and the branch-free exp; both keep the entry within ~1 ulp of the reference's formula.
__global__ void __launch_bounds__(256, 3) cov_build_g0_kernel(const __grid_constant__ KernelSpec spec,
                                                           const double* __restrict__ Xs, int N,
                                                           const double* __restrict__ noise, double* __restrict__ K) {
  extern __shared__ double sm[];
  const int dim = spec.dim;
  double* Xr = sm;               // [dim][TR]   coordinate-major: a lane's 4 rows are 32 contiguous bytes
  double* Xc = sm + TR * dim;    // [dim][TCW]  (a warp shares its 4 columns: broadcast reads)
  // linear block index -> (tile row tr, tile col tc) of the lower-triangular tile grid: tc*TCW <= tr*TR + TR - 1
  constexpr int kColsPerRow = TR / TCW;  // col tiles that fit under one row tile's diagonal extent
  int tr = static_cast<int>((sqrt(8.0 * (blockIdx.x / kColsPerRow) + 1.0) - 1.0) * 0.5);
  while (tr * (tr + 1) / 2 * kColsPerRow > static_cast<int>(blockIdx.x)) --tr;
  while ((tr + 1) * (tr + 2) / 2 * kColsPerRow <= static_cast<int>(blockIdx.x)) ++tr;
  const int tc = blockIdx.x - tr * (tr + 1) / 2 * kColsPerRow;
  const int row0 = tr * TR, colbase = tc * TCW;
  if (colbase >= N) return;
  // coalesced, asynchronous (LDGSTS) slab loads, transposed on the fly into the coordinate-major shared layout
  for (int e = threadIdx.x; e < TR * dim; e += blockDim.x) {
    const int r = e / dim, k = e - r * dim;
    const bool ok = row0 + r < N;
    cp_async8(Xr + k * TR + r, ok ? Xs + static_cast<size_t>(row0) * dim + e : Xs, ok);
  }
  for (int e = threadIdx.x; e < TCW * dim; e += blockDim.x) {
    const int c = e / dim, k = e - c * dim;
    const bool ok = colbase + c < N;
    cp_async8(Xc + k * TCW + c, ok ? Xs + static_cast<size_t>(colbase) * dim + e : Xs, ok);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  for (int sub = 0; sub < TCW / TC; ++sub) {
  const int col0 = colbase + sub * TC;
  if (col0 > row0 + TR - 1 || col0 >= N) break;  // sub-tile entirely above the diagonal
  // 2-D register blocking inside a warp: lane = (lr, lc) owns rows roff..roff+3 and cols coff..coff+3 of a 32x16 warp
  // tile, so an LDS.128 of the row slab touches 8 distinct 16-byte chunks (1 wavefront) instead of 32 (4 wavefronts)
  // and the kernel is FP64- rather than shared-memory-bound.  8 warps = 4 (rows) x 2 (cols) cover 128 x 32.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int roff4 = ((warp & 3) * 32 + (lane & 7) * 4) / 4;   // in units of 4 rows
  const int coff4 = ((warp >> 2) * 16 + (lane >> 3) * 4) / 4; // in units of 4 cols
  const int rg = roff4, cg = coff4;
  const double nz = noise[0];
  const bool se = spec.kernel == CMOE_KERNEL_SQUARE_EXPONENTIAL;
  // -r^2/2 = x_i.x_j - |x_i|^2/2 - |x_j|^2/2 : one FMA per dimension and entry (the absolute error of the expanded
  // form is ~1e-16 * (|x_i|^2 + |x_j|^2), i.e. a relative perturbation of k of that size — far below the noise term)
  double t[4][4], hc[4] = {0.0, 0.0, 0.0, 0.0};
  double hr[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k = 0; k < dim; ++k) {
    const double2 ra = *reinterpret_cast<const double2*>(Xr + k * TR + rg * 4);
    const double2 rb = *reinterpret_cast<const double2*>(Xr + k * TR + rg * 4 + 2);
    hr[0] = fma(-0.5 * ra.x, ra.x, hr[0]);
    hr[1] = fma(-0.5 * ra.y, ra.y, hr[1]);
    hr[2] = fma(-0.5 * rb.x, rb.x, hr[2]);
    hr[3] = fma(-0.5 * rb.y, rb.y, hr[3]);
  }
#pragma unroll
  for (int rr = 0; rr < 4; ++rr)
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) t[rr][cc] = 0.0;
#pragma unroll 2
  for (int k = 0; k < dim; ++k) {
    const double2 ra = *reinterpret_cast<const double2*>(Xr + k * TR + rg * 4);
    const double2 rb = *reinterpret_cast<const double2*>(Xr + k * TR + rg * 4 + 2);
    const double2 ca = *reinterpret_cast<const double2*>(Xc + k * TCW + sub * TC + cg * 4);
    const double2 cb = *reinterpret_cast<const double2*>(Xc + k * TCW + sub * TC + cg * 4 + 2);
    const double xr[4] = {ra.x, ra.y, rb.x, rb.y};
    const double xc[4] = {ca.x, ca.y, cb.x, cb.y};
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      hc[rr] = fma(-0.5 * xc[rr], xc[rr], hc[rr]);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) t[rr][cc] = fma(xr[rr], xc[cc], t[rr][cc]);
    }
  }
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int gc = col0 + cg * 4 + cc;
    if (gc >= N) continue;
    double v[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const double e = t[rr][cc] + (hr[rr] + hc[cc]);  // = -r^2/2
      if (se) {
        v[rr] = spec.alpha * exp_fast(e);
      } else {
        const double r2 = fmax(0.0, -2.0 * e);
        const double arg = kSqrt5 * sqrt(r2);
        v[rr] = spec.alpha * exp_fast(-arg) * (1.0 + arg + 5.0 / 3.0 * r2);
      }
      if (row0 + rg * 4 + rr == gc) v[rr] = spec.alpha + nz;  // exact diagonal: k(x, x) = alpha
    }
    const int gr = row0 + rg * 4;
    double* dst = K + static_cast<size_t>(gc) * N + gr;
    if (gr + 3 < N && (N % 2 == 0)) {
      // 16-byte aligned when N is even (gr is a multiple of 4)
      reinterpret_cast<double2*>(dst)[0] = make_double2(v[0], v[1]);
      reinterpret_cast<double2*>(dst)[1] = make_double2(v[2], v[3]);
    } else {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
        if (gr + rr < N) dst[rr] = v[rr];
    }
  }
  }  // sub-tiles
}

// generic path (derivative observations): one thread per point pair, writes the (1+g)x(1+g) block
__global__ void __launch_bounds__(256) cov_build_generic_kernel(const __grid_constant__ KernelSpec spec,
                                                                const double* __restrict__ X, int N,
                                                                const double* __restrict__ noise,
                                                                double* __restrict__ K) {
  const int dim = spec.dim, b = 1 + spec.g, n = N * b;
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);  // row point (fast)
  const int j = blockIdx.y * 8 + (threadIdx.x >> 5);   // col point
  if (i >= N || j >= N || i < j) return;
  const double* p1 = X + static_cast<size_t>(i) * dim;
  const double* p2 = X + static_cast<size_t>(j) * dim;
  const KParts kp = kernel_parts(spec, weighted_sqdist(spec, p1, p2));
  for (int cn = 0; cn < b; ++cn) {
    const int a2 = cn ? spec.derivs[cn - 1] : -1;
    for (int m = 0; m < b; ++m) {
      const int a1 = m ? spec.derivs[m - 1] : -1;
      const int row = i * b + m, col = j * b + cn;
      if (row >= col) {
        double v = cov_entry(spec, kp, p1, p2, a1, a2);
        if (row == col) v += noise[m];
        K[static_cast<size_t>(col) * n + row] = v;
      }
    }
  }
}

// K(X, P): thread per (sampled point i, P point j), writes the (1+g)x(1+gs) block; rows fastest.
__global__ void __launch_bounds__(256) mix_cov_kernel(const __grid_constant__ KernelSpec spec,
                                                      const double* __restrict__ X, int N,
                                                      const double* __restrict__ P, int num,
                                                      const int* __restrict__ dPs, int gs,
                                                      double* __restrict__ out) {
  const int dim = spec.dim, b = 1 + spec.g, bs = 1 + gs, n = N * b;
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const int j = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (i >= N || j >= num) return;
  const double* p1 = X + static_cast<size_t>(i) * dim;
  const double* p2 = P + static_cast<size_t>(j) * dim;
  const KParts kp = kernel_parts(spec, weighted_sqdist(spec, p1, p2));
  for (int cn = 0; cn < bs; ++cn) {
    const int a2 = cn ? dPs[cn - 1] : -1;
    for (int m = 0; m < b; ++m) {
      const int a1 = m ? spec.derivs[m - 1] : -1;
      out[static_cast<size_t>(j * bs + cn) * n + i * b + m] = cov_entry(spec, kp, p1, p2, a1, a2);
    }
  }
}

__global__ void philox_table_kernel(uint64_t seed, uint64_t first_draw, int num_draws, int per_draw,
                                    double* __restrict__ out) {
  const int pairs = (per_draw + 1) / 2;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(num_draws) * pairs) return;
  const int i = static_cast<int>(idx / pairs), k = static_cast<int>(idx % pairs);
  double a, b;
  philox_normal_pair(seed, first_draw + i, k, a, b);
  out[static_cast<size_t>(i) * per_draw + 2 * k] = a;
  if (2 * k + 1 < per_draw) out[static_cast<size_t>(i) * per_draw + 2 * k + 1] = b;
}

}  // namespace

void build_covariance(const KernelSpec& spec, const double* X, const double* Xs, int N, const double* noise,
                      double* K, cudaStream_t s) {
  if (spec.g == 0) {
    const int trows = (N + TR - 1) / TR;
    dim3 grid(trows * (trows + 1) / 2 * (TR / TCW));
    const size_t smem = static_cast<size_t>(TR + TCW) * spec.dim * sizeof(double);
    cov_build_g0_kernel<<<grid, 256, smem, s>>>(spec, Xs, N, noise, K);
  } else {
    dim3 grid((N + 31) / 32, (N + 7) / 8);
    cov_build_generic_kernel<<<grid, 256, 0, s>>>(spec, X, N, noise, K);
  }
  count_launch();
  CMOE_CUDA(cudaGetLastError());
}

void build_mix_covariance(const KernelSpec& spec, const double* X, int N, const double* P, int num, const int* dPs,
                          int gs, double* out, cudaStream_t s) {
  if (num == 0 || N == 0) return;
  dim3 grid((N + 31) / 32, (num + 7) / 8);
  mix_cov_kernel<<<grid, 256, 0, s>>>(spec, X, N, P, num, dPs, gs, out);
  count_launch();
  CMOE_CUDA(cudaGetLastError());
}

void philox_normals_device(uint64_t seed, uint64_t first_draw, int num_draws, int per_draw, double* out,
                           cudaStream_t s) {
  const size_t total = static_cast<size_t>(num_draws) * ((per_draw + 1) / 2);
  if (!total) return;
  philox_table_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(seed, first_draw, num_draws, per_draw,
                                                                               out);
  count_launch();
  CMOE_CUDA(cudaGetLastError());
}

}  // namespace cmoe
