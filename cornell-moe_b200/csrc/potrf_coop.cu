// Blocked FP64 Cholesky for large systems: ONE cooperative launch per 256-column outer panel.
//
// Replaces ComputeCholeskyFactorL (reference gpp_linear_algebra.cpp:109-148; pivot test `> 1e-16`, failing leading
// minor k+1) for n >= 1024.  The launch-per-64-columns path of linalg.cu (8 dependent launches per outer panel, ~40 us
// per 64-column step) stays for small systems, where exact IEEE sqrt / divide keep the known-answer cases exact.
//
// Launch p does two things with the whole machine (grid = #SMs, 1 CTA / SM, all co-resident):
//   * the trailing update with panel p-1 (K = 256):  C -= L_rows L_cols^T on the FP64 tensor pipe (DMMA m8n8k4), operand
//     chunks streamed by TMA tensor copies (cp.async.bulk.tensor.2d -> SASS UTMALDG) through a 5-stage mbarrier ring,
//     accumulators initialised from C so the epilogue is a plain store.  Tiles come from one atomic counter; the
//     tiles that cover the columns of panel p ("(a)" tiles, 128 x 64, top rows first) are handed out before the rest
//     ("(b)" tiles, 128 x 128) and bump a per-row-tile arrival counter when they are done;
//   * the factorisation of panel p itself: CTA r < ceil((n - p0) / 64) owns 64 rows of the panel (a 64 x 256 slab that
//     stays in shared memory) and starts as soon as the (a) tiles of its rows have arrived.  The four 64-column steps
//     are chained by release/acquire flags through L2 instead of kernel boundaries: CTA kk factors the diagonal block
//     (register Cholesky of two 32x32 halves by one warp), publishes L_kk^-1, every other CTA forms
//     X = A_ik L_kk^-T as ONE 64x64x64 DMMA product (no per-row substitution recurrence) and applies it to its own
//     columns; the X tiles of the rows inside the panel are published the same way.  The critical path per 64 columns
//     is  factor + one flag hand-over + two 64^3 products.
// When its panel work is done a CTA joins the (b) tiles, so the trailing update of panel p-1 overlaps the (latency-
// bound) chain of panel p on the same launch — no side stream, no events.
//
// Determinism: every C tile receives exactly one update per panel, with a fixed K order; which CTA computes it does
// not matter.  Failure: the factoring CTA writes the 1-based leading-minor index to *flag; every wait loop polls it.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <ctime>

#include "device_math.cuh"
#include "internal.cuh"
#include "linalg_dev.cuh"
#include "ptx_util.cuh"

namespace cmoe {
namespace {

constexpr int CW = 256;                // outer panel width = K depth of the trailing update
constexpr int CTHREADS = 512;          // 16 warps
constexpr int TM = 128;                // rows of a trailing-update tile
constexpr int KC = 32;                 // K chunk per pipeline stage (16 with 5 stages measured 1.6 % slower: every chunk
                                       // boundary is an mbarrier wait + shared-memory latency that all warps meet together)
constexpr int OLD = 132;               // rows of an operand box = leading dimension in shared memory (4 mod 16)
constexpr int STAGES = 3;
constexpr int PREFETCH = 2;            // chunks in flight ahead of the consumers
constexpr int SLD = NB + 4;            // leading dimension of slab / staging blocks (68 = 4 mod 16)
constexpr int BLK = SLD * NB;          // doubles per 64-column block buffer (34 816 B, a multiple of 128)
constexpr int OPBOX = KC * OLD;        // doubles per operand box
constexpr int NBUF = 6;                // 4 slab blocks + 2 staging buffers
constexpr int kSyncPerPanel = 160;     // ints: [0] (a) counter, [1..4] F, [5..20] X[j][kk], [21] (b) counter, [22..25] arrivals of
                                       // the diagonal block's 64 x 64 tiles, [32..] row-tile arrivals
constexpr int kTraceSlots = 64;
constexpr size_t kCoopSmem = static_cast<size_t>(NBUF) * BLK * sizeof(double) + 128;
static_assert(STAGES * 2 * OPBOX <= NBUF * BLK, "operand ring must fit into the slab buffers");

struct CoopParams {
  double* A;
  int lda, n;
  int p0, pw;        // panel factored by this launch
  int* flag;         // failure index (device)
  int* sync;         // this launch's sync area
  double* scratch;   // [4][BLK] published inverses
  int* abort_flag;   // set when a wait gives up
  int gate;          // >= 0: CTAs of the rows below the diagonal block keep taking trailing tiles until diagonal block
                     // `gate` has been factored (see chol_step_kernel); -1: they join the chain at once
  unsigned long long* trace;  // optional [gridDim][kTraceSlots] time stamps (CMOE_CHOL_TRACE), else NULL
};

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// returns false if the factorisation failed / aborted while waiting
__device__ __forceinline__ bool wait_flag(const int* f, int target, const int* fail, int* abort_flag) {
  for (unsigned spins = 0;; ++spins) {
    if (ld_acquire(f) >= target) return true;
    if ((spins & 31u) == 31u && (ld_volatile(fail) != 0 || ld_volatile(abort_flag) != 0)) return false;
    if (spins > (1u << 24)) {
      atomicExch(abort_flag, 1);
      return false;
    }
    __nanosleep(32);
  }
}

__device__ __forceinline__ void trace_mark(const CoopParams& P, int slot) {
  if (P.trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.trace[static_cast<size_t>(blockIdx.x) * kTraceSlots + slot] = t;
    P.trace[static_cast<size_t>(blockIdx.x) * kTraceSlots + kTraceSlots - 1] = slot;  // last checkpoint reached
  }
}

__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// --------------------------------------------------------------------------------------------------------------
// Trailing-update tile:  C(128 x TN) -= A[row0.., q0..q0+255] * A[col0.., q0..q0+255]^T
// --------------------------------------------------------------------------------------------------------------
template <int TMR, int TN>
__device__ __forceinline__ void gemm_tile(const CUtensorMap* mapOp, double* ring, uint64_t* full, uint64_t* empty,
                                          uint32_t& gchunk, double* __restrict__ A, int lda, int n, int row0, int col0,
                                          int q0) {
  constexpr int MF = TMR / 32;  // 8-row fragments per warp (4 x 4 warps)
  constexpr int NF = TN / 32;   // 8-column fragments per warp
  constexpr int NCH = CW / KC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp & 3) * (TMR / 4), wn = (warp >> 2) * (TN / 4);
  const int lr = lane >> 2, lc = lane & 3;
  const uint32_t g0 = gchunk;
  auto issue = [&](int c) {
    const uint32_t g = g0 + c;
    const int st = g % STAGES;
    mbar_wait_spin(&empty[st], ((g / STAGES) + 1) & 1);
    mbar_expect_tx(&full[st], 2 * OPBOX * sizeof(double));
    tma_load_2d(ring + st * 2 * OPBOX, mapOp, row0, q0 + c * KC, &full[st]);
    tma_load_2d(ring + st * 2 * OPBOX + OPBOX, mapOp, col0, q0 + c * KC, &full[st]);
  };
  if (tid == 0) {
#pragma unroll 1
    for (int c = 0; c < PREFETCH; ++c) issue(c);
  }
  // accumulators start as C: D = (-A) B + C, so the epilogue is a plain store
  double acc[MF][NF][2];
  double* Cg = A + static_cast<size_t>(col0) * lda + row0;
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    const int r = wm + i * 8 + lr;
#pragma unroll
    for (int j = 0; j < NF; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = wn + j * 8 + lc * 2 + h;
        acc[i][j][h] = (row0 + r < n && col0 + c < n) ? Cg[static_cast<size_t>(c) * lda + r] : 0.0;
      }
  }
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    if (tid == 0 && c + PREFETCH < NCH) issue(c + PREFETCH);
    const uint32_t g = g0 + c;
    const int st = g % STAGES;
    mbar_wait_spin(&full[st], (g / STAGES) & 1);
    const double* as = ring + st * 2 * OPBOX;
    const double* bs = as + OPBOX;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      double a[MF], b[NF];
#pragma unroll
      for (int i = 0; i < MF; ++i) a[i] = -as[(kk + lc) * OLD + wm + i * 8 + lr];
#pragma unroll
      for (int j = 0; j < NF; ++j) b[j] = bs[(kk + lc) * OLD + wn + j * 8 + lr];
#pragma unroll
      for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
  gchunk = g0 + NCH;
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    const int r = wm + i * 8 + lr;
#pragma unroll
    for (int j = 0; j < NF; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = wn + j * 8 + lc * 2 + h;
        if (row0 + r < n && col0 + c < n) Cg[static_cast<size_t>(c) * lda + r] = acc[i][j][h];
      }
  }
}

// --------------------------------------------------------------------------------------------------------------
// 64 x 64 tile product out of shared memory for the panel chain: acc (+)= (+-A)(64 x K) * B(64 x K)^T, operands stored
// [k][SLD].  16 warps x (16 x 16).  TRI: B is the transposed inverse of a lower-triangular block (B[k][c] = 0 for
// k > c), so the K loop of a warp stops at its last column.
// --------------------------------------------------------------------------------------------------------------
// Warp -> 16 x 16 sub-tile.  Standard: warp w -> (rows 16 (w % 4), columns 16 (w / 4)).  LOWER (symmetric update of a
// diagonal block: only the lower triangle is ever read): the ten lower sub-tiles are dealt to warps 0..9 in an order that
// puts at most three on a scheduler (the standard map leaves four on one), warps 10..15 idle (wm < 0).
struct WarpTile {
  int wm, wn;
};
__device__ __forceinline__ WarpTile warp_tile(bool lower) {
  const int warp = threadIdx.x >> 5;
  if (!lower) return {(warp & 3) * 16, (warp >> 2) * 16};
  // (row, col) in units of 16, in the order 00 10 11 20 21 22 30 31 32 33
  if (warp >= 10) return {-1, -1};
  const int row = (warp >= 6) ? 3 : (warp >= 3) ? 2 : (warp >= 1) ? 1 : 0;
  return {row * 16, (warp - row * (row + 1) / 2) * 16};
}

template <bool NEG, bool TRI>
__device__ __forceinline__ void tile64_mma(double (&acc)[2][2][2], const double* As, const double* Bs,
                                           const WarpTile wt) {
  const int lane = threadIdx.x & 31;
  const int wm = wt.wm, wn = wt.wn;
  const int lr = lane >> 2, lc = lane & 3;
  const int kend = TRI ? wn + 16 : NB;
  if (wm < 0) return;
  // two independent accumulator sets (even / odd k-steps): the product is a chain of dependent DMMAs per accumulator
  // and this routine sits on the panel's critical path — half the chain, summed at the end
  double acc2[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc2[i][j][0] = acc2[i][j][1] = 0.0;
#pragma unroll 2
  for (int kk = 0; kk < kend; kk += 8) {
    double a[2], b[2], a2[2], b2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double v = As[(kk + lc) * SLD + wm + i * 8 + lr];
      const double v2 = As[(kk + 4 + lc) * SLD + wm + i * 8 + lr];
      a[i] = NEG ? -v : v;
      a2[i] = NEG ? -v2 : v2;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      b[j] = Bs[(kk + lc) * SLD + wn + j * 8 + lr];
      b2[j] = Bs[(kk + 4 + lc) * SLD + wn + j * 8 + lr];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        dmma_m8n8k4(acc2[i][j][0], acc2[i][j][1], a2[i], b2[j]);
      }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      acc[i][j][0] += acc2[i][j][0];
      acc[i][j][1] += acc2[i][j][1];
    }
}

// fragment <-> block coordinates of tile64_mma
#define CMOE_FRAG_LOOP(WT, BODY)                                        \
  if ((WT).wm >= 0) {                                                   \
    const int lane_ = threadIdx.x & 31;                                 \
    const int wm_ = (WT).wm, wn_ = (WT).wn;                             \
    const int lr_ = lane_ >> 2, lc_ = lane_ & 3;                        \
    _Pragma("unroll") for (int i = 0; i < 2; ++i)                       \
    _Pragma("unroll") for (int j = 0; j < 2; ++j)                       \
    _Pragma("unroll") for (int h = 0; h < 2; ++h) {                     \
      const int row = wm_ + i * 8 + lr_, col = wn_ + j * 8 + lc_ * 2 + h; \
      BODY                                                              \
    }                                                                   \
  }

// NOTE: none of the shared-memory pointers below may be __restrict__: they carry data BETWEEN threads across barriers
// (a restrict-qualified `sfail` let the compiler keep the first half's value in a register across __syncthreads(); the
// warp that wrote the failure then returned alone and the CTA dead-locked on its next barrier).

// A follower: x <- x L^-T over the 32 columns [cbase, cbase + 32) of the factor under construction, four columns at a
// time as they are announced by chol32_warp_pair.
template <typename Emit>
__device__ __forceinline__ void solve_follow(double (&x)[32], double* LT, double* rd, int cbase, volatile int* prog,
                                             Emit&& emit) {
#pragma unroll 1
  for (int kb = 0; kb < 32; kb += 4) {
    while (*prog < kb + 4) __nanosleep(64);  // a tight poll would queue shared-memory loads in front of the factoring warp's
    __threadfence_block();
    solve_steps_rot<32, 4, 4, true>(x, LT, rd, cbase + kb, emit);
  }
}

// One 8 x 8 block of a 32 x 32 product on the FP64 tensor pipe: acc (+)= sum_{k0 <= k < k1} A(r, k) B(c, k) for the block
// rows r = 8 bi .. and columns c = 8 bj ..; a_at / b_at fetch single operand entries (shared memory, any layout).  The
// SYRK / Y / Z pieces of the diagonal-block factorisation were per-thread dot products before: 96 shared-memory loads
// per thread and ~1.8 us each on the chain, bound by the load pipe; as 16 warp blocks they are 16 loads + 8 DMMA per warp.
template <typename FA, typename FB>
__device__ __forceinline__ void mma_block8(double (&acc)[2], int bi, int bj, int k0, int k1, FA&& a_at, FB&& b_at) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
  for (int k = k0; k < k1; k += 4) dmma_m8n8k4(acc[0], acc[1], a_at(bi * 8 + lr, k + lc), b_at(bj * 8 + lr, k + lc));
}

// Factor the 64 x 64 diagonal block held in `blk` ([c*SLD + r], lower part meaningful) and build the packet
// M[m*SLD + j] = (L^-1)[j][m] in `Minv`.  LT (transposed factor, [c*LTS + r]) lives in `LT`.  Returns 0 or the 1-based
// index of the failing pivot.  All CTHREADS threads call.  Warp 0 factors; warp 1 (inverse of the diagonal 32-blocks)
// and warp 2 (rows 32..63 of the first half) trail it column by column.
__device__ __forceinline__ int factor_diag64(double* blk, double* LT, double* Minv, double* cb, double* rd,
                                             volatile int* sfail, volatile int* prog, const CoopParams& P, int tbase) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Nothing but the column buffer needs clearing: LT is only read where it has been written or into window slots that are
  // never emitted (any finite or non-finite garbage is discarded there); of Minv only the lower-triangular part of the
  // packet is produced, read (Y, Z) and published — the other half of the packet is zero in global memory for good.
  for (int e = tid; e < 256; e += CTHREADS) cb[e] = 0.0;
  if (tid == 0) *prog = 0;
  __syncthreads();
  // the X tile this CTA stored as a consumer of the previous step is announced from here by a warp that idles through
  // the first half anyway (every thread's stores precede the barrier above; barrier + one thread's fence publishes them)
  // (block kk >= 1 is factored by the CTA that consumed step kk - 1 as row block kk: its X flag is sync[5 + 4 kk + kk - 1])
  const int kk_self = (tbase - 32) / 6;
  if (kk_self > 0 && tid == 3 * 32) {
    __threadfence();
    fence_proxy_async();
    st_release(P.sync + 5 + kk_self * 4 + (kk_self - 1), 1);
  }
  const int tfine = tbase - 24;  // the consumer slots of this step are free in the factoring CTA
  trace_mark(P, tfine + 0);
  if (warp == 0) {
    double a[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? blk[c * SLD + lane] : 0.0;
    const int f = chol32_warp_pipe(a, lane, cb, LT, rd, 0, prog);
    if (lane == 0) *sfail = f;
    trace_mark(P, tfine + 1);
  } else if (warp == 1) {
    // column `lane` of L11^-1 = row `lane` of L11^-T: e_lane L11^-T
    double x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = (c == lane) ? 1.0 : 0.0;
    solve_follow(x, LT, rd, 0, prog, [&](int k, double v) { Minv[lane * SLD + k] = v; });
  } else if (warp == 2) {
    // L21 = A21 L11^-T: lane r solves row 32 + r
    double x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = blk[c * SLD + 32 + lane];
    solve_follow(x, LT, rd, 0, prog, [&](int k, double v) {
      blk[k * SLD + 32 + lane] = v;
      LT[k * LTS + 32 + lane] = v;
    });
  }
  __syncthreads();
  trace_mark(P, tbase + 1);
  if (*sfail) return *sfail;
  {
    // A22 -= L21 L21^T (lower blocks) and Y = L21 * L11^-1 (parked in the lower-left block of Minv): warp w -> block (w / 4, w % 4)
    const int bi = warp >> 2, bj = warp & 3, lr = lane >> 2, lc = lane & 3;
    if (bi >= bj) {
      double acc[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) acc[h] = blk[(32 + bj * 8 + lc * 2 + h) * SLD + 32 + bi * 8 + lr];
      mma_block8(acc, bi, bj, 0, 32, [&](int r, int k) { return -blk[k * SLD + 32 + r]; },
                 [&](int c, int k) { return blk[k * SLD + 32 + c]; });
#pragma unroll
      for (int h = 0; h < 2; ++h) blk[(32 + bj * 8 + lc * 2 + h) * SLD + 32 + bi * 8 + lr] = acc[h];
    }
    {
      // Y[i][t] = sum_{m >= t} L21[i][m] I11[m][t]
      double acc[2] = {0.0, 0.0};
      mma_block8(acc, bi, bj, bj * 8, 32, [&](int i, int m) { return LT[m * LTS + 32 + i]; },
                 [&](int t, int m) { return Minv[t * SLD + m]; });
#pragma unroll
      for (int h = 0; h < 2; ++h) Minv[(bj * 8 + lc * 2 + h) * SLD + 32 + bi * 8 + lr] = acc[h];
    }
    if (tid == 0) *prog = 0;
  }
  trace_mark(P, tfine + 2);
  __syncthreads();
  trace_mark(P, tbase + 2);
  if (warp == 0) {
    double a[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? blk[(32 + c) * SLD + 32 + lane] : 0.0;
    const int f = chol32_warp_pipe(a, lane, cb, LT, rd, 32, prog);
    if (lane == 0) *sfail = f ? 32 + f : 0;
    trace_mark(P, tfine + 3);
  } else if (warp == 1) {
    double x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = (c == lane) ? 1.0 : 0.0;
    solve_follow(x, LT, rd, 32, prog, [&](int k, double v) { Minv[(32 + lane) * SLD + k] = v; });
  }
  __syncthreads();
  trace_mark(P, tbase + 3);
  if (*sfail) return *sfail;
  {
    // Z = -L22^-1 * Y: Z[i][t] = -sum_{m <= i} Inv22[i][m] Y[m][t], in place over Y (all blocks read before any is written)
    const int bi = warp >> 2, bj = warp & 3, lr = lane >> 2, lc = lane & 3;
    double acc[2] = {0.0, 0.0};
    mma_block8(acc, bi, bj, 0, bi * 8 + 8, [&](int i, int m) { return -Minv[(32 + m) * SLD + 32 + i]; },
               [&](int t, int m) { return Minv[t * SLD + 32 + m]; });
    trace_mark(P, tfine + 4);
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) Minv[(bj * 8 + lc * 2 + h) * SLD + 32 + bi * 8 + lr] = acc[h];
  }
  __syncthreads();
  trace_mark(P, tbase + 4);
  return 0;
}

// --------------------------------------------------------------------------------------------------------------
// Panel chain of row block r (64 rows) of the panel [p0, p0 + pw).
// --------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void panel_role(const CUtensorMap* mapBlk, const CoopParams& P, double* buf, uint64_t* pbar,
                                           uint32_t& pphase, double* colbuf, double* rd, volatile int* sfail, volatile int* sflag,
                                       volatile int* prog, int r) {
  // out of line: tell the compiler what the caller knew (shared-memory pointers -> LDS / STS instead of generic LD / ST)
  __builtin_assume(__isShared(buf));
  __builtin_assume(__isShared(colbuf));
  __builtin_assume(__isShared(rd));
  __builtin_assume(__isShared(pbar));
  __builtin_assume(__isShared(const_cast<const int*>(sfail)));
  __builtin_assume(__isShared(const_cast<const int*>(sflag)));
  __builtin_assume(__isShared(const_cast<const int*>(prog)));
  const int tid = threadIdx.x;
  const int nkk = (P.pw + NB - 1) / NB;
  const int rows0 = P.p0 + NB * r;
  const int jmax = min(r, nkk - 1);
  int* F = P.sync + 1;
  int* X = P.sync + 5;
  int* rowready = P.sync + 32;
  auto bcast_wait = [&](const int* f, int target) -> bool {
    if (tid == 0) *sflag = wait_flag(f, target, P.flag, P.abort_flag) ? 1 : 0;
    __syncthreads();
    const bool ok = *sflag != 0;
    __syncthreads();
    return ok;
  };
  trace_mark(P, 1);
  if (P.p0 > 0) {
    // rows of the diagonal block: their own 64 x 64 tiles; rows below: the 128 x 64 tiles of their row pair
    const int* ready = (r < 4) ? P.sync + 22 + r : rowready + (r >> 1);
    const int expect = (r < 4) ? min(r, nkk - 1) + 1 : nkk;
    if (!bcast_wait(ready, expect)) return;
  }
  if (tid == 0) {
    fence_proxy_async();
    mbar_expect_tx(pbar, static_cast<uint32_t>((jmax + 1) * BLK * sizeof(double)));
    for (int j = 0; j <= jmax; ++j) tma_load_2d(buf + j * BLK, mapBlk, rows0, P.p0 + NB * j, pbar);
  }
  trace_mark(P, 2);
  mbar_wait_spin(pbar, pphase);
  pphase ^= 1;
  trace_mark(P, 3);
  double* stageM = buf + 4 * BLK;
  double* stageL = buf + 5 * BLK;
  for (int kk = 0; kk < nkk; ++kk) {
    const int k0 = P.p0 + NB * kk;
    const int nb = min(NB, P.n - k0);
    double* blkk = buf + kk * BLK;
    if (r == kk) {
      // ---- factor the diagonal block, publish L_kk^-1 ----
      if (nb < NB) {
        for (int c = nb + tid; c < NB; c += CTHREADS) blkk[c * SLD + c] = 1.0;  // identity padding (rows are zero-filled)
      }
      __syncthreads();
      trace_mark(P, 32 + kk * 6 + 0);
      const int failed = factor_diag64(blkk, stageM, stageL, colbuf, rd, sfail, prog, P, 32 + kk * 6);
      if (failed) {
        if (tid == 0) {
          *P.flag = k0 + failed;
          __threadfence();
        }
        return;
      }
      // publish the inverse first (the chain waits for it), the factor tile itself afterwards.  Packet entry
      // [m * SLD + j] = (L^-1)[j][m]: only j >= m is non-zero and only that half is ever written (the buffer was cleared
      // when it was allocated)
      double* G = P.scratch + kk * BLK;
      for (int e = tid; e < BLK; e += CTHREADS) {
        const int m = e / SLD, j = e - m * SLD;
        if (j >= m && j < NB) G[e] = stageL[e];
      }
      trace_mark(P, 8 + kk * 6 + 5);
      __syncthreads();
      if (tid == 0) {
        // barrier, then ONE thread's fence and release: cumulative over the block's stores that precede the barrier (the
        // pattern of a cooperative-groups grid barrier) — the other 511 threads do not wait for their stores to drain
        __threadfence();
        fence_proxy_async();
        st_release(F + kk, 1);
      }
      double* Ab = P.A + static_cast<size_t>(k0) * P.lda + k0;
      for (int e = tid; e < NB * NB; e += CTHREADS) {
        const int rr = e & (NB - 1), c = e >> 6;
        if (rr >= c && rr < nb && c < nb) Ab[static_cast<size_t>(c) * P.lda + rr] = stageM[c * LTS + rr];
      }
      trace_mark(P, 32 + kk * 6 + 5);
      return;
    }
    // ---- X = A_rk L_kk^-T ----
    trace_mark(P, 8 + kk * 6 + 0);
    if (!bcast_wait(F + kk, 1)) return;
    trace_mark(P, 8 + kk * 6 + 1);
    if (tid == 0) {
      fence_proxy_async();
      mbar_expect_tx(pbar, BLK * sizeof(double));
      tma_bulk_g2s(stageM, P.scratch + kk * BLK, BLK * sizeof(double), pbar);
    }
    mbar_wait_spin(pbar, pphase);
    pphase ^= 1;
    trace_mark(P, 8 + kk * 6 + 2);
    {
      double acc[2][2][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      const WarpTile wt = warp_tile(false);
      tile64_mma<false, true>(acc, blkk, stageM, wt);
      __syncthreads();  // every warp has read its A rows
      double* Ag = P.A + static_cast<size_t>(k0) * P.lda + rows0;
      CMOE_FRAG_LOOP(wt, {
        blkk[col * SLD + row] = acc[i][j][h];
        if (rows0 + row < P.n && col < nb) Ag[static_cast<size_t>(col) * P.lda + row] = acc[i][j][h];
      })
      __syncthreads();  // X (shared-memory copy) visible to every warp: it is an operand of the updates below
      trace_mark(P, 8 + kk * 6 + 3);
    }
    // the X tiles of the rows inside the panel are operands of everybody's updates: publish them (stores are already in
    // flight).  The CTA that factors next (r == kk + 1) first updates its own diagonal block, so the fence drains in
    // the shadow of that product instead of on the chain.
    bool published = !(r < nkk);
    auto publish = [&]() {
      if (!published) {
        __threadfence();
        __syncthreads();
        if (tid == 0) {
          fence_proxy_async();
          st_release(X + r * 4 + kk, 1);
        }
        published = true;
      }
    };
    if (r != kk + 1) publish();
    // ---- apply to the columns to the right (inside the panel) ----
    for (int jb = kk + 1; jb <= jmax; ++jb) {
      const double* Bs = blkk;
      if (jb != r) {
        if (!bcast_wait(X + jb * 4 + kk, 1)) return;
        if (tid == 0) {
          fence_proxy_async();
          mbar_expect_tx(pbar, BLK * sizeof(double));
          tma_load_2d(stageL, mapBlk, P.p0 + NB * jb, k0, pbar);
        }
        mbar_wait_spin(pbar, pphase);
        pphase ^= 1;
        Bs = stageL;
      }
      double* blkj = buf + jb * BLK;
      double acc[2][2][2];
      WarpTile wt = warp_tile(false);
      // own diagonal tile: only its lower triangle is ever read.  (The balanced map warp_tile(true) would save ~0.5 us
      // here but costs the pivot loop of factor_diag64 a spilled loop-carried register with this register budget.)
      if (jb == r && wt.wn > wt.wm) wt.wm = -1;
      CMOE_FRAG_LOOP(wt, { acc[i][j][h] = blkj[col * SLD + row]; })
      tile64_mma<true, false>(acc, blkk, Bs, wt);
      CMOE_FRAG_LOOP(wt, { blkj[col * SLD + row] = acc[i][j][h]; })
      __syncthreads();  // stageL is reused by the next column block; blkj complete before it becomes an operand
      trace_mark(P, 8 + kk * 6 + 4);
    }
    if (r == kk + 1 && r < nkk) published = true;  // this CTA factors next: factor_diag64 releases the flag
    publish();
  }
  trace_mark(P, 60);
}

__global__ void __launch_bounds__(CTHREADS, 1)
    chol_step_kernel(const __grid_constant__ CUtensorMap mapOp, const __grid_constant__ CUtensorMap mapBlk,
                     const __grid_constant__ CoopParams P) {
  extern __shared__ __align__(128) unsigned char coop_smem_raw[];
  double* buf = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(coop_smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  // the round-up hides the address space from the compiler: without this every operand fragment load of the trailing
  // update is a generic LD.E instead of LDS (seen in the SASS of the DMMA loop)
  __builtin_assume(__isShared(buf));
  __shared__ __align__(8) uint64_t full[STAGES];
  __shared__ __align__(8) uint64_t empty[STAGES];
  __shared__ __align__(8) uint64_t pbar;
  __shared__ double colbuf[256];
  __shared__ double rd[NB];
  __shared__ int s_tile, sfail, sflag, prog;
  trace_mark(P, 0);
  if (ld_volatile(P.flag) != 0) return;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CTHREADS / 32);
    }
    mbar_init(&pbar, 1);
    fence_proxy_async();
  }
  __syncthreads();
  const int n = P.n, p0 = P.p0;
  const int nrb = (n - p0 + NB - 1) / NB;        // panel row blocks
  // CTA r < nrb owns row block r of the panel.  A row block below the diagonal block spends most of the chain waiting
  // for the next diagonal block (its own share of a step is a few microseconds of a ~30 us step).  While the trailing
  // update is the bottleneck of the launch (P.gate >= 0) such a CTA therefore keeps taking trailing tiles until diagonal
  // block `gate` is done and runs through its steps afterwards, every earlier flag already set.
  const bool is_p = static_cast<int>(blockIdx.x) < nrb;
  const bool gated = is_p && P.gate >= 0 && static_cast<int>(blockIdx.x) >= 4;
  bool p_done = false;
  uint32_t gchunk = 0, pphase = 0;
  // trailing update with the previous panel
  int num_a = 0, total = 0, nct_a = 0, nbt = 0;
  const int q0 = p0 - CW;
  const int tb0 = p0 + CW;  // origin of the (b) tile grid
  // (a) tiles = update of panel p's own columns.  The rows of the panel's diagonal block (the first four 64-row blocks,
  // whose CTAs carry the chain) are cut into 64 x 64 tiles that are handed out FIRST: the chain starts after ~half the
  // time of a regular 128 x 64 tile.  Row block r needs its lower tiles (r, 0..min(r, nct_a-1)).
  int num_small = 0, small_rows = 0;
  if (p0 > 0) {
    const int nrt = (n - p0 + TM - 1) / TM;
    nct_a = (P.pw + NB - 1) / NB;
    small_rows = min(4, nrb);
    for (int r = 0; r < small_rows; ++r) num_small += min(r, nct_a - 1) + 1;
    num_a = num_small + max(0, nrt - 2) * nct_a;
    if (n > tb0) {
      nbt = (n - tb0 + TM - 1) / TM;
      total = num_a + nbt * (nbt + 1) / 2;
    } else {
      total = num_a;
    }
  }
  // two tile counters: a CTA with pending panel work only ever claims (a) tiles — claiming a (b) tile before the chain
  // would park that tile for the whole chain (seen in the trace: +40 us per panel)
  int* counter_a = P.sync;
  int* counter_b = P.sync + 21;
  int* rowready = P.sync + 32;
  // the four chain CTAs never pick up (a) tiles: nothing may delay the start of the chain
  bool a_done = (num_a == 0) || static_cast<int>(blockIdx.x) < small_rows;
  int* small_ready = P.sync + 22;
  const int* gate_flag = P.sync + 1 + max(0, P.gate);  // F flag of diagonal block `gate`
  bool b_left = true;  // thread 0's view of the (b) tile pool
  for (;;) {
    if (tid == 0) {
      int t = -2;
      if (!a_done) {
        t = atomicAdd(counter_a, 1);
        if (t >= num_a) t = -3;  // (a) tiles exhausted
      }
      if (t < 0) {
        const bool hold_rows = gated && !p_done && b_left && ld_acquire(gate_flag) == 0;
        if (!is_p || p_done || hold_rows) {
          const int tb = num_a + atomicAdd(counter_b, 1);
          if (tb < total) {
            t = tb;
          } else {
            b_left = false;
            t = (is_p && !p_done) ? -3 : total;  // nothing left to fill the wait with: join the chain / leave
          }
        }
      }
      s_tile = t;
    }
    __syncthreads();
    const int t = s_tile;
    __syncthreads();
    if (t < 0) {
      a_done = true;
      if (is_p && !p_done) {
        panel_role(&mapBlk, P, buf, &pbar, pphase, colbuf, rd, &sfail, &sflag, &prog, blockIdx.x);
        p_done = true;
        __syncthreads();
      }
      continue;
    }
    if (t >= num_a) a_done = true;
    if (t >= total) break;
    trace_mark(P, t < num_a ? 4 : 61);
    if (t < num_small) {
      int r = 0, rem = t;
      while (rem >= min(r, nct_a - 1) + 1) {
        rem -= min(r, nct_a - 1) + 1;
        ++r;
      }
      gemm_tile<64, 64>(&mapOp, buf, full, empty, gchunk, P.A, P.lda, n, p0 + r * NB, p0 + rem * NB, q0);
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicAdd(small_ready + r, 1);
      trace_mark(P, 5);
    } else if (t < num_a) {
      const int ti = 2 + (t - num_small) / nct_a, tj = (t - num_small) % nct_a;
      gemm_tile<128, 64>(&mapOp, buf, full, empty, gchunk, P.A, P.lda, n, p0 + ti * TM, p0 + tj * NB, q0);
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicAdd(rowready + ti, 1);
      trace_mark(P, 5);
    } else {
      int rem = t - num_a, tj = 0;
      while (rem >= nbt - tj) {
        rem -= nbt - tj;
        ++tj;
      }
      const int ti = tj + rem;
      gemm_tile<128, 128>(&mapOp, buf, full, empty, gchunk, P.A, P.lda, n, tb0 + ti * TM, tb0 + tj * TM, q0);
    }
    trace_mark(P, 6);
  }
  trace_mark(P, 62);
}

PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

}  // namespace

bool make_tensor_map_2d(CUtensorMap* map, const double* base, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols) {
  auto enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {rows, cols};
  const cuuint64_t strides[1] = {ld * sizeof(double)};
  const cuuint32_t box[2] = {box_rows, box_cols};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS;
}

// Returns false when the cooperative path does not apply (small / odd n, more row blocks than SMs, no cooperative
// launch): nothing has been touched in that case and the caller takes the launch-per-step path.
bool potrf_lower_coop(double* A, int n, int* flag, cudaStream_t s) {
  if (legacy_linalg() || n < 1024 || (n & 1) || (reinterpret_cast<uintptr_t>(A) & 15)) return false;
  int dev = 0, sms = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop || (n + NB - 1) / NB > sms) return false;
  CUtensorMap mapOp, mapBlk;
  if (!make_tensor_map_2d(&mapOp, A, n, n, n, OLD, KC) || !make_tensor_map_2d(&mapBlk, A, n, n, n, SLD, NB)) return false;
  CMOE_CUDA(cudaFuncSetAttribute(chol_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(kCoopSmem)));
  const int npanels = (n + CW - 1) / CW;
  // workspace kept per host thread and device: cudaMalloc / cudaFree inside the timed factorisation cost up to
  // milliseconds on a loaded host (seen as 3.4 -> 5.6 ms on one box with identical device-side spans)
  struct Workspace {
    int device = -1;
    DevBuf<int> sync;
    DevBuf<double> scratch;
  };
  static thread_local Workspace ws;
  if (ws.device != dev) {
    ws.sync.release();
    ws.scratch.release();
    ws.device = dev;
  }
  ws.sync.ensure(static_cast<size_t>(npanels) * kSyncPerPanel + 1);
  if (ws.scratch.count < static_cast<size_t>(4) * BLK) {
    ws.scratch.alloc(static_cast<size_t>(4) * BLK);
    CMOE_CUDA(cudaMemsetAsync(ws.scratch.p, 0, ws.scratch.count * sizeof(double), s));  // the packets' zero halves
  }
  DevBuf<int>& sync = ws.sync;
  DevBuf<double>& scratch = ws.scratch;
  CMOE_CUDA(cudaMemsetAsync(sync.p, 0, (static_cast<size_t>(npanels) * kSyncPerPanel + 1) * sizeof(int), s));
  CMOE_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
  int* abort_flag = sync.p + static_cast<size_t>(npanels) * kSyncPerPanel;
  // CMOE_CHOL_TRACE=<file>: %globaltimer stamps of thread 0 of every CTA at the role / step boundaries of every panel
  const char* trace_path = std::getenv("CMOE_CHOL_TRACE");
  DevBuf<unsigned long long> trace;
  const size_t per_panel = static_cast<size_t>(sms) * kTraceSlots;
  if (trace_path) {
    trace.alloc(per_panel * npanels);
    CMOE_CUDA(cudaMemsetAsync(trace.p, 0, trace.count * sizeof(unsigned long long), s));
  }
  // trailing row tiles (128 rows) from which a panel counts as update-bound (the chain takes ~130 us, a 128 x 128 tile
  // ~40 us on one SM) and the diagonal block the lower rows wait for; CMOE_CHOL_GATE="<tiles>,<block>" overrides (a huge
  // tile count disables)
  int gate_min_tiles = 22, gate_block = 1;
  if (const char* e = std::getenv("CMOE_CHOL_GATE")) std::sscanf(e, "%d,%d", &gate_min_tiles, &gate_block);
  gate_block = std::max(0, std::min(gate_block, 3));
  for (int p0 = 0, pi = 0; p0 < n; p0 += CW, ++pi) {
    // hold the rows below the diagonal block back while the trailing update, not the chain, bounds the launch
    const int nbt = n > p0 + CW ? (n - p0 - CW + TM - 1) / TM : 0;
    const int gate = (p0 > 0 && p0 + CW < n && nbt >= gate_min_tiles) ? gate_block : -1;
    CoopParams P{A, n, n, p0, std::min(CW, n - p0), flag, sync.p + static_cast<size_t>(pi) * kSyncPerPanel, scratch.p,
                 abort_flag, gate, trace_path ? trace.p + per_panel * pi : nullptr};
    void* args[] = {&mapOp, &mapBlk, &P};
    CMOE_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(chol_step_kernel), dim3(sms), dim3(CTHREADS), args,
                                          kCoopSmem, s));
    count_launch();
  }
  if (trace_path) {
    // watchdog: if the launches have not drained after 20 s, dump where every CTA is and keep waiting
    cudaStream_t side;
    CMOE_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    std::vector<unsigned long long> ht(trace.count);
    bool done = false;
    for (int it = 0; it < 2000 && !(done = cudaStreamQuery(s) == cudaSuccess); ++it) {
      struct timespec ts = {0, 10 * 1000 * 1000};
      nanosleep(&ts, nullptr);
    }
    CMOE_CUDA(cudaMemcpyAsync(ht.data(), trace.p, trace.count * sizeof(unsigned long long), cudaMemcpyDeviceToHost, side));
    CMOE_CUDA(cudaStreamSynchronize(side));
    cudaStreamDestroy(side);
    if (FILE* f = std::fopen(trace_path, "w")) {
      std::fprintf(f, "# n=%d done=%d; per panel and CTA: last checkpoint, then stamps (ns, relative to the CTA's slot 0)\n", n,
                   done ? 1 : 0);
      for (int pi = 0; pi < npanels; ++pi)
        for (int c = 0; c < sms; ++c) {
          const unsigned long long* t = ht.data() + per_panel * pi + static_cast<size_t>(c) * kTraceSlots;
          if (t[0] == 0) continue;
          std::fprintf(f, "p%d c%d last=%llu :", pi, c, t[kTraceSlots - 1]);
          for (int k = 1; k < kTraceSlots - 1; ++k)
            if (t[k]) std::fprintf(f, " %d:%lld[%llx]", k, static_cast<long long>(t[k] - t[0]), t[k]);
          std::fprintf(f, "\n");
        }
      std::fclose(f);
    }
  }
  int aborted = 0;
  CMOE_CUDA(cudaMemcpyAsync(&aborted, abort_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  CMOE_CUDA(cudaStreamSynchronize(s));
  CMOE_REQUIRE(!aborted, CMOE_ERR_RUNTIME,
               "cooperative Cholesky: a dependency wait timed out (device shared with another long-running kernel?)");
  return true;
}

}  // namespace cmoe
