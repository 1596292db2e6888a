// Device-side building blocks of the dense FP64 factorisation shared by linalg.cu (launch-per-step path, small
// systems) and potrf_coop.cu (one cooperative launch per outer panel): warp-level register Cholesky of a 32x32 block,
// rotating-window triangular substitution, the FP64 tensor-pipe instruction.
#pragma once

#include "internal.cuh"

namespace cmoe {
namespace {

constexpr int NB = 64;        // Cholesky block size == K depth of one DMMA tile product
constexpr int LDT = NB + 4;   // smem leading dimension: 68 = 4 (mod 16) makes the DMMA fragment loads conflict-free
constexpr double kPivotTol = 1.0e-16;  // gpp_linear_algebra.cpp:118

// --------------------------------------------------------------------------------------------------------------
// Diagonal-block factorisation + panel solve.  The column recurrence sqrt -> scale -> update is the latency-critical
// chain of the whole factorisation (n/64 dependent block steps), so it runs out of REGISTERS: a warp factors a 32x32
// block with lane r holding row r and the pivot / multiplier columns travelling by shuffle — no barriers, no shared
// memory round trips inside the recurrence.
// --------------------------------------------------------------------------------------------------------------
constexpr int PT = 128;  // threads of the panel kernel: one panel row per thread

// The register-resident recurrences below are written as SHORT LOOPS over a rotating register window (after step k
// the window is shifted so that slot 0 is always the pivot column) instead of fully unrolled triangles: straight-line
// code of thousands of instructions that execute once each runs at instruction-fetch speed (measured: 18 us for a
// 64x64 factorisation, 25 us for the panel solve), a loop body of a few hundred instructions stays in the I-cache.
// Slots that rotate past the end of the row compute on padding and are never stored.

// In-register Cholesky of a 32x32 block: on entry a[c] = A[lane][c] (c <= lane meaningful).  The finished column j of
// L is written to LT[(col0 + j) * LTS + col0 + lane] (transposed factor) and rd[col0 + j] = 1/L_jj (FAST) or L_jj.
// Returns 0 or the 1-based index of the first pivot that fails `> 1e-16` (gpp_linear_algebra.cpp:118,141-142); the
// outcome is warp-uniform and only evaluated at the end (the arithmetic after a failed pivot is discarded).
// The finished column is broadcast through a double-buffered shared column (one STS + one __syncwarp per column).
// FAST: sqrt and divide through one reciprocal square root + Newton corrections — the same results as sqrt()/"/" to
// the last bit in all but rare halfway cases (and exactly when the true results are representable) at a third of the
// dependent latency; small systems (known-answer cases) keep IEEE sqrt / divide.
constexpr int LTS = NB + 2;  // row stride of the transposed factor (even: 16-byte aligned rows)
template <bool FAST>
__device__ __forceinline__ int chol32_warp(double (&a)[32], int lane, double* colbuf /* [2][64], [32..63] = 0 */,
                                           double* LT, double* rd, int col0) {
  int fail = 0;
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    const double piv = __shfl_sync(0xffffffffu, a[0], j);
    fail = (fail == 0 && !(piv > kPivotTol)) ? j + 1 : fail;
    double l, q;
    if (FAST) {
      const double y = rsqrt(piv);
      l = piv * y;
      l = fma(0.5 * y, fma(-l, l, piv), l);
      q = a[0] * y;
      q = fma(fma(-q, l, a[0]), y, q);
    } else {
      l = sqrt(piv);
      q = a[0] / l;
    }
    const double lj = (lane == j) ? l : q;
    if (lane >= j) LT[(col0 + j) * LTS + col0 + lane] = lj;
    if (FAST) {
      double r = rsqrt(piv);            // ~ 1/l
      r = fma(fma(-l, r, 1.0), r, r);   // one Newton step on 1/l: no divide on the per-column critical path
      if (lane == j) rd[col0 + j] = r;
    } else {
      if (lane == j) rd[col0 + j] = l;
    }
    double* col = colbuf + (j & 1) * 64;
    col[lane] = lj;
    __syncwarp();
    // window slot k <-> column j + k; multipliers L[j+k][j] = col[j+k] (zero beyond the block)
#pragma unroll
    for (int k = 1; k < 32; ++k) a[k - 1] = fma(-lj, col[j + k], a[k]);
    a[31] = 0.0;
  }
  return fail;
}

// sqrt(p) and 1/sqrt(p) for a normal p > 0: hardware seed (MUFU.RSQ64H) + two coupled Goldschmidt steps: one MUFU and
// five dependent FP64 operations (measured on B200: dependent DFMA 11 cycles, seed 13) instead of the library rsqrt
// (66 cycles) plus a Newton correction (33) on the pivot chain.  Results within ~1 ulp.
__device__ __forceinline__ void sqrt_rsqrt(double p, double& l, double& y) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(p));
  double g = p * y0, h = 0.5 * y0;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  l = fma(g, r, g);
  y = 2.0 * fma(h, r, h);
}

// Same factorisation, two columns per round: the 2 x 2 pivot block [l11 0; l21 l22] is formed from three shuffles, both
// multiplier columns travel through ONE shared-memory round trip and the window is updated by a rank-2 step.
// LEAN = false: the arithmetic (operation order and roundings) is exactly that of two successive rounds of
// chol32_warp<true>.  LEAN = true (the cooperative large-n path): pivots through sqrt_rsqrt, multipliers as a * (1/l)
// without the correction step — the column recurrence is the latency-critical chain of the whole factorisation
// (measured per column, one warp alone on the SM: 324 cycles single-column, 270 paired, see microbench_chain.cu).
// colbuf: [2][2][64] with [32..63] of every column zero.
// prog (optional, shared memory): columns finished so far, for follower warps that trail the factorisation.  The
// column data (LT, rd) and the counter are shared-memory stores of the same warp, issued in program order behind a
// __syncwarp(); no fence is placed on the pivot chain (measured: a block fence by one lane costs ~70 cycles there).
template <bool LEAN = false>
__device__ __forceinline__ int chol32_warp_pair(double (&a)[32], int lane, double* colbuf,
                                                double* LT, double* rd, int col0, volatile int* prog = nullptr) {
  int fail = 0;
#pragma unroll 1
  for (int j = 0; j < 32; j += 2) {
    const double p11 = __shfl_sync(0xffffffffu, a[0], j);
    const double p21 = __shfl_sync(0xffffffffu, a[0], j + 1);
    const double p22 = __shfl_sync(0xffffffffu, a[1], j + 1);
    fail = (fail == 0 && !(p11 > kPivotTol)) ? j + 1 : fail;
    if (fail) break;  // warp-uniform: nothing after a failed pivot is used; do not feed 0 / NaN to rsqrt
    double y1, l11, q1, l21;
    if (LEAN) {
      sqrt_rsqrt(p11, l11, y1);
      q1 = a[0] * y1;
      l21 = p21 * y1;
    } else {
      y1 = rsqrt(p11);
      l11 = p11 * y1;
      l11 = fma(0.5 * y1, fma(-l11, l11, p11), l11);
      q1 = a[0] * y1;
      q1 = fma(fma(-q1, l11, a[0]), y1, q1);
      l21 = p21 * y1;  // what lane j+1 computes as its q1 (same operations on the same operands)
      l21 = fma(fma(-l21, l11, p21), y1, l21);
    }
    // column j+1 after the update by column j
    const double d2 = fma(-l21, l21, p22);
    fail = (fail == 0 && !(d2 > kPivotTol)) ? j + 2 : fail;
    if (fail) break;
    const double a1 = fma(-q1, l21, a[1]);
    double y2, l22, q2;
    if (LEAN) {
      sqrt_rsqrt(d2, l22, y2);
      q2 = a1 * y2;
    } else {
      y2 = rsqrt(d2);
      l22 = d2 * y2;
      l22 = fma(0.5 * y2, fma(-l22, l22, d2), l22);
      q2 = a1 * y2;
      q2 = fma(fma(-q2, l22, a1), y2, q2);
    }
    const double lj1 = (lane == j) ? l11 : q1;
    const double lj2 = (lane == j + 1) ? l22 : q2;
    if (lane >= j) LT[(col0 + j) * LTS + col0 + lane] = lj1;
    if (lane >= j + 1) LT[(col0 + j + 1) * LTS + col0 + lane] = lj2;
    if (LEAN) {
      if (lane == j) rd[col0 + j] = y1;
      if (lane == j + 1) rd[col0 + j + 1] = y2;
    } else {
      if (lane == j) rd[col0 + j] = fma(fma(-l11, y1, 1.0), y1, y1);  // one Newton step on 1/l from y ~ 1/l
      if (lane == j + 1) rd[col0 + j + 1] = fma(fma(-l22, y2, 1.0), y2, y2);
    }
    double* c1 = colbuf + ((j >> 1) & 1) * 128;
    double* c2 = c1 + 64;
    c1[lane] = lj1;
    c2[lane] = lj2;
    __syncwarp();
    if (prog && lane == 0) *prog = j + 2;
    // window slot k <-> column j + k
#pragma unroll
    for (int k = 2; k < 32; ++k) a[k - 2] = fma(-lj2, c2[j + k], fma(-lj1, c1[j + k], a[k]));
    a[30] = 0.0;
    a[31] = 0.0;
  }
  if (prog && fail && lane == 0) *prog = 1 << 20;  // release the followers (their results are discarded)
  return fail;
}

// Software-pipelined form of the paired factorisation (lean pivots).  Measured: in chol32_warp_pair half of a round is
// the ISSUE time of the 60 window FMAs + 60 shared-memory loads that sit, in program order, between one round's pivots
// and the next round's shuffles (an in-order warp cannot start the next dependent chain before they have issued).
// Here only the two window slots the next pivots need are updated right away; the rest of round j's rank-2 update is
// carried as "pending" (multipliers in registers, columns in the other parity buffer) and executed inside round j+1,
// in one branch-free basic block together with that round's dependent pivot chain, so the scheduler fills the chain's
// latency gaps with it.  A failed pivot is recorded, not branched on (the raw RSQ64H seed has no slow path).
__device__ __forceinline__ int chol32_warp_pipe(double (&a)[32], int lane, double* colbuf, double* LT, double* rd,
                                                int col0, volatile int* prog = nullptr) {
  int fail = 0;
  double plj1 = 0.0, plj2 = 0.0;            // pending multipliers of the previous round (zero: nothing pending)
  const double* pc1 = colbuf + 128;         // previous round's column buffers (start: zeros)
  const double* pc2 = colbuf + 192;
  double p11 = __shfl_sync(0xffffffffu, a[0], 0);
  double p21 = __shfl_sync(0xffffffffu, a[0], 1);
  double p22 = __shfl_sync(0xffffffffu, a[1], 1);
#pragma unroll 1
  for (int j = 0; j < 32; j += 2) {
    // ---- pivot chain of columns j, j+1 (dependent) ----
    fail = (fail == 0 && !(p11 > kPivotTol)) ? j + 1 : fail;
    double y1, l11, y2, l22;
    sqrt_rsqrt(p11, l11, y1);
    const double q1 = a[0] * y1;
    const double l21 = p21 * y1;
    const double d2 = fma(-l21, l21, p22);
    fail = (fail == 0 && !(d2 > kPivotTol)) ? j + 2 : fail;
    sqrt_rsqrt(d2, l22, y2);
    const double a1 = fma(-q1, l21, a[1]);
    const double q2 = a1 * y2;
    const double lj1 = (lane == j) ? l11 : q1;
    const double lj2 = (lane == j + 1) ? l22 : q2;
    // ---- pending rank-2 update of the previous round on slots 2..31 (independent of the chain above) ----
#pragma unroll
    for (int k = 2; k < 32; ++k) a[k] = fma(-plj2, pc2[j + k], fma(-plj1, pc1[j + k], a[k]));
    // ---- publish the two columns ----
    if (lane >= j) LT[(col0 + j) * LTS + col0 + lane] = lj1;
    if (lane >= j + 1) LT[(col0 + j + 1) * LTS + col0 + lane] = lj2;
    if (lane == j) rd[col0 + j] = y1;
    if (lane == j + 1) rd[col0 + j + 1] = y2;
    double* c1 = colbuf + ((j >> 1) & 1) * 128;
    double* c2 = c1 + 64;
    c1[lane] = lj1;
    c2[lane] = lj2;
    __syncwarp();
    if (prog && lane == 0) *prog = j + 2;
    // ---- the two slots the next pivots need, then their shuffles ----
    const double n0 = fma(-lj2, c2[j + 2], fma(-lj1, c1[j + 2], a[2]));
    const double n1 = fma(-lj2, c2[j + 3], fma(-lj1, c1[j + 3], a[3]));
    p11 = __shfl_sync(0xffffffffu, n0, (j + 2) & 31);
    p21 = __shfl_sync(0xffffffffu, n0, (j + 3) & 31);
    p22 = __shfl_sync(0xffffffffu, n1, (j + 3) & 31);
    // window: slot k <-> column j + 2 + k; slots 2.. still lack this round's update (done in the next round)
    a[0] = n0;
    a[1] = n1;
#pragma unroll
    for (int k = 4; k < 32; ++k) a[k - 2] = a[k];
    a[30] = 0.0;
    a[31] = 0.0;
    plj1 = lj1;
    plj2 = lj2;
    pc1 = c1;
    pc2 = c2;
  }
  if (prog && fail && lane == 0) *prog = 1 << 20;
  return fail;
}

// STEPS steps of  x <- x L^-T  for one row held in a rotating register window of W slots: slot i holds column
// cbase + i on entry and column cbase + STEPS + i on exit.  The solved value of column k is handed to emit(k, value).
// LT[k*LTS + c] = L[c][k]; cbase and UNR are multiples of 2 and LTS is even, so pairs of multipliers come as one
// 16-byte broadcast load; rows of LT must be followed by readable padding (slots past the end of the row read it and
// are never emitted).
template <int W, int STEPS, int UNR, bool FAST, typename Emit>
__device__ __forceinline__ void solve_steps_rot(double (&x)[W], const double* LT,
                                                const double* rd, int cbase, Emit&& emit) {
  static_assert(STEPS % UNR == 0 && UNR % 2 == 0 && W % 2 == 0, "even windows, whole bodies");
#pragma unroll 1
  for (int kb = 0; kb < STEPS; kb += UNR) {
#pragma unroll
    for (int s = 0; s < UNR; ++s) {
      const int k = cbase + kb + s;
      const double xk = FAST ? x[s] * rd[k] : x[s] / rd[k];
      emit(k, xk);
      const double* lt = LT + k * LTS + cbase + kb;  // 16-byte aligned
      if (((s + 1) & 1) != 0) x[s + 1] = fma(-xk, lt[s + 1], x[s + 1]);
#pragma unroll
      for (int i = (s + 2) & ~1; i < W; i += 2) {
        const double2 v = *reinterpret_cast<const double2*>(lt + i);
        x[i] = fma(-xk, v.x, x[i]);
        x[i + 1] = fma(-xk, v.y, x[i + 1]);
      }
    }
#pragma unroll
    for (int i = 0; i + UNR < W; ++i) x[i] = x[i + UNR];
#pragma unroll
    for (int i = W - UNR; i < W; ++i) x[i] = 0.0;
  }
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

}  // namespace
}  // namespace cmoe
