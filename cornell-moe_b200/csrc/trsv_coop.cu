// Single-launch triangular solve for ONE right-hand side on a large factor (the K^-1 y of a big fit):
// TriangularMatrixVectorSolve / CholeskyFactorLMatrixVectorSolve, reference gpp_linear_algebra.cpp:160-193, .hpp:220.
//
// L is cut into 128-wide block rows; x_b = L_bb^-1 (rhs_b - sum_{j<b} L_bj x_j) (forward; the transposed solve mirrors
// it).  Every block row has one OWNER CTA and S-1 HELPER CTAs, all co-resident (cooperative launch, 1 CTA / SM):
//   * helpers accumulate the products with the blocks x_j that are already published (release/acquire flags through
//     L2), each over a strided share of j, and hand their partial sums to the owner — the factor is streamed once,
//     spread over ~120 SMs instead of the 40 of the first version, whose bottom CTAs could not keep up with the chain;
//   * the owner inverts its 128 x 128 diagonal block in shared memory while it waits (in-place blocked inversion:
//     four 32 x 32 register inversions, then the 64- and 128-level off-diagonal blocks), prefetches the tile of the
//     LAST dependency into registers, and, once that x_j arrives, needs one 128 x 128 register product, the partial sums
//     and one 128 x 128 product with the inverse: ~2 us per block on the critical path instead of ~13 us.
// Hand-overs carry no flags: the solution vector and the helpers' partial sums start out as a sentinel bit pattern
// (all ones — a NaN no arithmetic produces) and every consumer thread polls the very word it needs (L1-bypassing loads)
// until it is a number.  One L2 round trip per hand-over instead of fence + flag store + flag poll + data load.
// Fixed combination order (helper 0, 1, ..., then the owner's tile): results are bit-identical run to run.
#include "device_math.cuh"
#include "internal.cuh"
#include "ptx_util.cuh"

namespace cmoe {
namespace {

constexpr int VB = 128;    // unknowns per block row
constexpr int CT = 512;    // threads
constexpr int kMaxHelpers = 3;

__device__ __forceinline__ int ldv(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

constexpr unsigned long long kSentinel = ~0ull;  // the host fills hand-over buffers with 0xFF bytes

// spins until *p holds a number; false when the bound is hit or another CTA gave up (~1 s: far beyond any wait)
__device__ __forceinline__ bool poll_value(const double* p, double& v, int* abort_flag) {
  const volatile unsigned long long* q = reinterpret_cast<const volatile unsigned long long*>(p);
  for (unsigned spins = 0;; ++spins) {
    const unsigned long long bits = *q;
    if (bits != kSentinel) {
      v = __longlong_as_double(static_cast<long long>(bits));
      return true;
    }
    if ((spins & 255u) == 255u && ldv(abort_flag) != 0) return false;
    if (spins > (1u << 23)) {
      atomicExch(abort_flag, 1);
      return false;
    }
  }
}

// In-place inversion of the lower-triangular 128 x 128 block M ([c*VB + r], r >= c meaningful, upper part ignored and
// left untouched).  Rows / columns >= nb are treated as identity.  All CT threads call.
__device__ __forceinline__ void invert_lower128(double* M) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // level 0: the four 32 x 32 diagonal blocks, warp w < 4, lane t = column t of the inverse
  if (warp < 4) {
    double* D = M + (warp * 32) * VB + warp * 32;
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int m = 0; m < i; ++m) s = fma(-D[m * VB + i], x[m], s);  // broadcast read of L[i][m]
      x[i] = s / D[i * VB + i];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i >= lane) D[lane * VB + i] = x[i];
  }
  __syncthreads();
  // level 1 (blocks of 64: off-diagonal 32 x 32) and level 2 (the 64 x 64 off-diagonal block): Z = -I22 * L21 * I11
#pragma unroll 1
  for (int h = 32; h <= 64; h *= 2) {
    const int nblk = VB / (2 * h);          // 2 blocks of 64, then 1 block of 128
    const int per = h * h;                  // entries of one off-diagonal block
    const int total = nblk * per;           // 2048, then 4096
    constexpr int kMaxPer = 4096 / CT;      // entries per thread (8)
    double y[kMaxPer];
    // Y = L21 * I11  (I11 lower: sum over m >= t)
#pragma unroll
    for (int u = 0; u < kMaxPer; ++u) {
      const int e = tid + u * CT;
      y[u] = 0.0;
      if (e < total) {
        const int blk = e / per, i = e % h, t = (e % per) / h;  // entry (i, t) of block blk; i fastest
        const int o = blk * 2 * h;
        double s = 0.0;
        for (int m = t; m < h; ++m) s = fma(M[(o + m) * VB + o + h + i], M[(o + t) * VB + o + m], s);
        y[u] = s;
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kMaxPer; ++u) {
      const int e = tid + u * CT;
      if (e < total) {
        const int blk = e / per, i = e % h, t = (e % per) / h;
        const int o = blk * 2 * h;
        M[(o + t) * VB + o + h + i] = y[u];
      }
    }
    __syncthreads();
    // Z = -I22 * Y  (I22 lower: sum over m <= i)
#pragma unroll
    for (int u = 0; u < kMaxPer; ++u) {
      const int e = tid + u * CT;
      y[u] = 0.0;
      if (e < total) {
        const int blk = e / per, i = e % h, t = (e % per) / h;
        const int o = blk * 2 * h;
        double s = 0.0;
        for (int m = 0; m <= i; ++m) s = fma(M[(o + h + m) * VB + o + h + i], M[(o + t) * VB + o + h + m], s);
        y[u] = -s;
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kMaxPer; ++u) {
      const int e = tid + u * CT;
      if (e < total) {
        const int blk = e / per, i = e % h, t = (e % per) / h;
        const int o = blk * 2 * h;
        M[(o + t) * VB + o + h + i] = y[u];
      }
    }
    __syncthreads();
  }
}

struct TrsvParams {
  const double* L;
  int n;
  const double* rhs;
  double* x;         // out (distinct from rhs), sentinel until its block is solved
  double* partial;   // [nblk][kMaxHelpers][VB], sentinel until a helper delivers
  int* abort_flag;
  int nblk, S;       // CTAs per block row (1 owner + S-1 helpers)
};

// product of one 128 x 128 tile with 128 entries of x, result (128 values) into out[] (shared or global):
//   forward  (TRANS = false): tile = L[b0.., j0..], out[r] = sum_c L[b0+r][j0+c] xj[c]
//   backward (TRANS = true):  tile = L[j0.., b0..], out[c] = sum_r L[j0+r][b0+c] xj[r]
// Loads are issued by load_tile() (no dependence on x), the arithmetic by apply_tile().
template <bool TRANS>
__device__ __forceinline__ void load_tile(double (&v)[32], const double* __restrict__ L, int n, int b0, int nb, int j0,
                                          int jn) {
  const int t = threadIdx.x;
  if (!TRANS) {
    const int r = t & (VB - 1), q = t >> 7;  // row r, column quarter q: coalesced down the columns
    const double* col = L + static_cast<size_t>(j0 + q * 32) * n + b0 + r;
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = (r < nb && q * 32 + c < jn) ? __ldcg(col + static_cast<size_t>(c) * n) : 0.0;
  } else {
    const int lane = t & 31, warp = t >> 5;  // warp owns columns 8w..8w+7, lanes run down the rows (contiguous)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = warp * 8 + i;
      const double* col = L + static_cast<size_t>(b0 + c) * n + j0;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[i * 4 + k] = (c < nb && lane + 32 * k < jn) ? __ldcg(col + lane + 32 * k) : 0.0;
    }
  }
}

// adds the product into acc (per-thread partials; reduce with reduce_acc)
template <bool TRANS>
__device__ __forceinline__ void apply_tile(const double (&v)[32], const double* xs, double (&acc)[8]) {
  const int t = threadIdx.x;
  if (!TRANS) {
    const int q = t >> 7;
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c & 7] = fma(v[c], xs[q * 32 + c], acc[c & 7]);
  } else {
    const int lane = t & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i] = fma(v[i * 4 + k], xs[lane + 32 * k], acc[i]);
  }
}

// combine the per-thread partials into out[0..127] (shared memory, valid after the trailing barrier)
template <bool TRANS>
__device__ __forceinline__ void reduce_acc(const double (&acc)[8], double* part /* [4][VB] */,
                                           double* out) {
  const int t = threadIdx.x;
  if (!TRANS) {
    const int r = t & (VB - 1), q = t >> 7;
    part[q * VB + r] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    __syncthreads();
    if (t < VB) out[t] = (part[t] + part[VB + t]) + (part[2 * VB + t] + part[3 * VB + t]);
  } else {
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double s = warp_sum(acc[i]);
      if (lane == 0) out[warp * 8 + i] = s;
    }
  }
  __syncthreads();
}

template <bool TRANS>
__global__ void __launch_bounds__(CT, 1) trsv_coop_kernel(const TrsvParams P) {
  extern __shared__ __align__(16) double Minv[];  // [VB][VB] owner only
  __shared__ double xs[VB];
  __shared__ double part[4 * VB];
  __shared__ double sum[VB];
  __shared__ double vv[VB];
  const int t = threadIdx.x;
  const int S = P.S;
  const int slot = static_cast<int>(blockIdx.x) / S, role = static_cast<int>(blockIdx.x) % S;  // role 0 = owner
  const int b = TRANS ? (P.nblk - 1 - slot) : slot;  // block row; slot = position in the dependency chain
  const int n = P.n, b0 = b * VB, nb = min(VB, n - b0);
  const int ndep = slot;                              // blocks this row depends on
  auto dep_block = [&](int k) { return TRANS ? (P.nblk - 1 - k) : k; };  // k-th dependency in publication order
  auto wait_x = [&](int blk) -> bool {  // block `blk` of the solution -> xs[], polled element-wise
    bool ok = true;
    if (t < VB) {
      double val = 0.0;
      if (blk * VB + t < n) ok = poll_value(P.x + blk * VB + t, val, P.abort_flag);
      xs[t] = val;
    }
    return __syncthreads_and(ok ? 1 : 0) != 0;
  };
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0;
  double v[32];

  if (role != 0) {
    // ---- helper: dependencies k = role-1, role-1 + (S-1), ... among the first ndep-1 (the last one is the owner's) ----
    const int H = S - 1;
    for (int k = role - 1; k < ndep - 1; k += H) {
      const int j = dep_block(k), j0 = j * VB, jn = min(VB, n - j0);
      load_tile<TRANS>(v, P.L, n, b0, nb, j0, jn);  // issued before the wait: the tile does not depend on x_j
      if (!wait_x(j)) return;
      apply_tile<TRANS>(v, xs, acc);
    }
    reduce_acc<TRANS>(acc, part, sum);
    if (t < VB) P.partial[(static_cast<size_t>(b) * kMaxHelpers + (role - 1)) * VB + t] = sum[t];
    return;
  }

  // ---- owner ----
  // diagonal block -> shared memory (identity padding), inverted in place while the chain is still far away
  for (int e = t; e < VB * VB; e += CT) {
    const int r = e & (VB - 1), c = e >> 7;
    double val = (r == c) ? 1.0 : 0.0;
    if (r < nb && c < nb && r >= c) val = P.L[static_cast<size_t>(b0 + c) * n + b0 + r];
    Minv[c * VB + r] = val;
  }
  __syncthreads();
  invert_lower128(Minv);
  const double my_rhs = (t < VB && t < nb) ? P.rhs[b0 + t] : 0.0;
  // with helpers the owner only multiplies the tile of the LAST dependency (prefetched: it does not depend on x);
  // without them (S == 1: more than #SMs / 2 block rows) it walks all dependencies itself.  The helpers' partial sums
  // depend on earlier blocks only: they are collected BEFORE the wait for the last x_j, off the chain.
  const bool last_only = (S > 1);
  if (ndep > 0 && last_only) {
    const int j = dep_block(ndep - 1), j0 = j * VB, jn = min(VB, n - j0);
    load_tile<TRANS>(v, P.L, n, b0, nb, j0, jn);
  }
  double base = my_rhs;
  {
    bool ok = true;
    if (t < VB && S > 1 && ndep > 1) {
      double tot = 0.0;
      for (int h = 0; h < S - 1 && ok; ++h) {
        double ph = 0.0;
        ok = poll_value(P.partial + (static_cast<size_t>(b) * kMaxHelpers + h) * VB + t, ph, P.abort_flag);
        tot += ph;
      }
      base = my_rhs - tot;
    }
    if (!__syncthreads_and(ok ? 1 : 0)) return;
  }
  if (last_only) {
    if (ndep > 0) {
      if (!wait_x(dep_block(ndep - 1))) return;
      apply_tile<TRANS>(v, xs, acc);
    }
  } else {
#pragma unroll 1
    for (int k = 0; k < ndep; ++k) {
      const int j = dep_block(k), j0 = j * VB, jn = min(VB, n - j0);
      load_tile<TRANS>(v, P.L, n, b0, nb, j0, jn);
      if (!wait_x(j)) return;
      apply_tile<TRANS>(v, xs, acc);
    }
  }
  reduce_acc<TRANS>(acc, part, sum);
  if (t < VB) vv[t] = base - sum[t];
  __syncthreads();
  // x_b = L_bb^-1 v (forward: sum over c <= r) or L_bb^-T v (backward: sum over r >= c)
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0;
  if (!TRANS) {
    const int r = t & (VB - 1), q = t >> 7;
#pragma unroll 8
    for (int c = q * 32; c < q * 32 + 32; ++c)
      if (c <= r) acc[c & 7] = fma(Minv[c * VB + r], vv[c], acc[c & 7]);
  } else {
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = warp * 8 + i;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = lane + 32 * k;
        if (r >= c) acc[i] = fma(Minv[c * VB + r], vv[r], acc[i]);
      }
    }
  }
  reduce_acc<TRANS>(acc, part, sum);
  if (t < nb) P.x[b0 + t] = sum[t];  // the store IS the signal
}

}  // namespace

namespace {

// workspace kept per host thread and device: cudaMalloc / cudaFree per call cost more than the solve itself
struct TrsvWorkspace {
  int device = -1;
  DevBuf<int> flags;
  DevBuf<double> out, partial;
};

// One or two chained solves (forward, then optionally the transposed one on its result) enqueued back to back: one set of
// sentinel fills, one abort read-back and one host synchronisation for the lot.
bool trsv_coop_run(const double* L, int n, double* x, bool first_trans, bool both, cudaStream_t s) {
  int dev = 0, sms = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  const int nblk = (n + VB - 1) / VB;
  if (!coop || nblk > sms) return false;
  const int S = std::max(1, std::min(1 + kMaxHelpers, sms / nblk));
  const size_t smem = static_cast<size_t>(VB) * VB * sizeof(double);
  CMOE_CUDA(cudaFuncSetAttribute(trsv_coop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  CMOE_CUDA(cudaFuncSetAttribute(trsv_coop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  static thread_local TrsvWorkspace ws;
  if (ws.device != dev) {
    ws.flags.release();
    ws.out.release();
    ws.partial.release();
    ws.device = dev;
  }
  const int nsolve = both ? 2 : 1;
  const size_t np = static_cast<size_t>(nblk) * kMaxHelpers * VB;
  if (ws.flags.count == 0) ws.flags.alloc(1);
  ws.out.ensure(static_cast<size_t>(2) * n);
  ws.partial.ensure(2 * np);
  CMOE_CUDA(cudaMemsetAsync(ws.flags.p, 0, sizeof(int), s));
  CMOE_CUDA(cudaMemsetAsync(ws.out.p, 0xFF, static_cast<size_t>(nsolve) * n * sizeof(double), s));  // sentinel
  CMOE_CUDA(cudaMemsetAsync(ws.partial.p, 0xFF, nsolve * np * sizeof(double), s));
  for (int k = 0; k < nsolve; ++k) {
    const bool trans = (k == 0) ? first_trans : true;
    TrsvParams P{L, n, k == 0 ? x : ws.out.p, ws.out.p + static_cast<size_t>(k) * n, ws.partial.p + k * np, ws.flags.p, nblk, S};
    void* args[] = {&P};
    void* fn = trans ? reinterpret_cast<void*>(trsv_coop_kernel<true>) : reinterpret_cast<void*>(trsv_coop_kernel<false>);
    CMOE_CUDA(cudaLaunchCooperativeKernel(fn, dim3(nblk * S), dim3(CT), args, smem, s));
    count_launch();
  }
  int aborted = 0;
  CMOE_CUDA(cudaMemcpyAsync(&aborted, ws.flags.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CMOE_CUDA(cudaStreamSynchronize(s));
  if (aborted) return false;
  CMOE_CUDA(cudaMemcpyAsync(x, ws.out.p + static_cast<size_t>(nsolve - 1) * n, static_cast<size_t>(n) * sizeof(double),
                            cudaMemcpyDeviceToDevice, s));
  CMOE_CUDA(cudaStreamSynchronize(s));  // the cached workspace may be reused from another stream of this thread
  return true;
}

}  // namespace

// false: not applicable (too many block rows for the device, no cooperative launch) or a wait gave up; x untouched
bool trsv_coop(const double* L, int n, double* x, bool trans, cudaStream_t s) {
  return trsv_coop_run(L, n, x, trans, false, s);
}

// x <- L^-T L^-1 x in one go (same contract)
bool trsv_coop_pair(const double* L, int n, double* x, cudaStream_t s) { return trsv_coop_run(L, n, x, false, true, s); }

}  // namespace cmoe
