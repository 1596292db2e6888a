// Fused q-KG Monte-Carlo kernels (the north-star hot path).
//
// One MC sample-evaluation of the reference = fantasise y = mu + L z at the q candidate points, re-solve K^-1 y on the
// (N+q) system, scan the discretisation set for the arg-min of the new posterior mean and run a line-search gradient
// descent from it (gpp_knowledge_gradient_optimization.cpp:87-113, 420-472; gpp_optimization.hpp:708-828).
// Here the re-solve is replaced by the algebraically identical rank-q update
//     mu+(x) = m + sum_j (beta_j - B_j . c) k(x, X_j) + sum_u c_u k(x, Xu_u),      c = L^-T z,  B = K^-1 K*
// so a sample touches no HBM at all: the per-candidate operands (scaled X, beta, B) are staged once per CTA into shared
// memory with TMA bulk copies and every lane runs one sample's whole inner optimisation out of registers.
//
// Mapping: one thread per MC sample; grid = (sample chunks, candidates).  Lanes that finish fetch the next sample of
// the chunk from a shared counter, so a warp only idles at the very end of a chunk although the line search has
// data-dependent trip counts.  The line search is a per-lane state machine around two warp-uniform routines:
// "evaluate mu+ and its gradient at my query point" (POINT round) and "evaluate all backtracking trials of this step
// along my gradient" (LINE round: eval_line / eval_line_gen for the SquareExponential kernel, eval_line_matern for
// Matern-5/2 without derivative observations); a warp vote picks the kind of round, see kg_mc_body.
// With derivative observations (kg_mc_gen_kernel) the per-sample weights are formed once per sample and streamed back
// through a per-lane cp.async ring (fill_*_weights, stream_point_weights).
#pragma once

#include <algorithm>

#include "device_math.cuh"
#include "internal.cuh"

#include "ptx_util.cuh"

namespace cmoe {

// Range contract of the shared-memory fast path.  Its exp (exp_tab, no per-call guard) needs |t| < 2^31 ln2 / 64 =
// 2.3e7, t = -|x~ - X~_j|^2 / 2 + ln(alpha).  The host only selects the fast path while every scaled operand (training,
// union, discrete points, inner-domain corners) lies within kFastPathRadius of the origin; a query point farther than
// kFarRadius is then at least kFarRadius - kFastPathRadius = 2000 length scales from every operand, so all its kernel
// values are exactly 0 and the evaluation is short-circuited; any other query point has |t| <= 6000^2 / 2 = 1.8e7.
#ifndef CMOE_MC_THREADS
#define CMOE_MC_THREADS 128
#endif
#ifndef CMOE_MC_MINBLOCKS
#define CMOE_MC_MINBLOCKS 3
#endif
constexpr int kMcThreads = CMOE_MC_THREADS;

constexpr double kFastPathRadius = 2000.0;
constexpr double kFarRadius = 4000.0;

struct KgMcParams {
  int N;        // training points
  int U;        // union points (= rows, g == 0)
  int ps;       // free coordinates (dim - num_fidelity)
  int dim;
  int M;        // discretisation set size (U + num_pts)
  int num_mc;
  int chunk;    // samples per CTA
  int use_smem;
  int g;          // derivative observations per point (0 on the fast path)
  int Q;          // rows of the union block = U * (1 + g)
  int pk_stride;  // doubles per training point in Pk: QP + 2 (g == 0) or 2 (g > 0: e_j only, the weights live in Wt)
  int derivs[8];  // observed partial-derivative indices (g <= 8 on the general path)
  int max_steps, max_restarts;
  double mean, mrc, tol, step_tol, alpha;
  const double* Xt;      // [N][DIM] scaled training points (zero padded)
  const double* Pk;      // [nc][N][QP+2]  (e_j, beta_j, B_j[0..QP))
  const double* Xu;      // [nc][U][DIM+2] (scaled union point, e_u, pad)
  const double* A;       // [nc][M][DIM]   unscaled start points (fidelity coords = 1, padding = 0)
  const double* recC;    // [nc][num_mc][QP]
  const int* recStart;   // [nc][num_mc]
  const double* alpha0;  // [max_steps]  pre_mult * (i+1)^-gamma
  double* outVal;        // [nc][num_mc]   -mu+(x*)  (the reference's best_function_value)
  double* outX;          // [nc][num_mc][DIM] scaled minimiser
  double* outH;          // [nc][num_mc]      -|scaled minimiser|^2 / 2 (consumed by kg_acc_kernel)
  unsigned long long* stats;  // [4]: posterior evaluations, accepted steps, point rounds, line batches (per lane)
  // general path (derivative observations): per-candidate weight table and the per-lane weight columns
  const double* Wt;  // [nc][1 + QP][N * (1+g)]: row 0 = beta~, row 1+u = B~[:, u]  (~: derivative rows divided by l_t)
  double* aw;        // [aw_slots][N * (1+g)][kMcThreads]: a = beta~ - B~ c of the sample each lane is working on
  int* work;         // work-item counter of the persistent grid (zeroed before the launch)
  int aw_slots;      // resident CTAs the scratch was sized for
  int chunks;        // work items per candidate
  int work_total;    // chunks * candidates of this launch
  int stage_ops;     // general path: scaled training points and e_j fit next to the ring in shared memory
  double lo[CMOE_MAX_DIM], hi[CMOE_MAX_DIM], inv_len[CMOE_MAX_DIM], len[CMOE_MAX_DIM];
};

struct KgAccParams {
  int N, U, dim, num_mc;
  int fast_exp;  // 1: the range contract of the table exp holds (see kFastPathRadius)
  int g;          // derivative observations per point (general path: rows = (N + U) * (1 + g))
  int derivs[8];
  double inv_len[CMOE_MAX_DIM];
  double alpha;
  const double* Xt;     // [N][DIM]
  const double* Xu;     // [nc][U][DIM+2]
  const double* recC;   // [nc][num_mc][QP]
  const double* outX;   // [nc][num_mc][DIM]
  const double* outH;   // [nc][num_mc]  -|x*|^2/2
  double* R;            // [nc][QP][N+U]   R[a][row] = sum_i c_ia k(row, x*_i)
  double* Gu;           // [nc][U][DIM]    sum_i c_iu Bpart(Xu_u, x*_i) x~*_i
  double* GkB;          // [nc][U]         sum_i c_iu Bpart(Xu_u, x*_i)
};

// exp(t) = 2^k * 2^(i/64) * e^r with |r| <= ln2/128: 64-entry table in shared memory + degree-5 Taylor polynomial
// (truncation r^6/720 <= 3.5e-17 relative).  10 FP64-pipe instructions instead of 15 for exp_fast; same integer-side
// clamp of the exponent, NO guard for |t| >= 2.3e7 (see kFastPathRadius).  Used by the shared-memory fast path of the
// fused q-KG kernel (+11 % throughput on the north-star shape); measured max relative error 3.0e-16 over [-690, 1]
// (profiles/exp_accuracy.py), exp(0) = 1 exactly.
__device__ const double kExp2Table[64] = {
#include "exp2_table64.inc"
};
__shared__ double g_exp_tab[64];
__device__ __forceinline__ void exp_table_stage() {
  if (threadIdx.x < 64) g_exp_tab[threadIdx.x] = kExp2Table[threadIdx.x];
  __syncthreads();
}
__device__ __forceinline__ double exp_tab(double t) {
  const double kShift = 6755399441055744.0;
  double nf = fma(t, 64.0 * 1.4426950408889634, kShift);
  const int n = __double2loint(nf);
  nf -= kShift;
  double r = fma(nf, -6.93147180369123816490e-01 / 64.0, t);
  r = fma(nf, -1.90821492927058770002e-10 / 64.0, r);
  double p = 8.3333333333333332e-03;
  p = fma(p, r, 4.1666666666666664e-02);
  p = fma(p, r, 1.6666666666666666e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= g_exp_tab[n & 63];
  const int k = max(n >> 6, -1000);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
template <bool TAB>
__device__ __forceinline__ double exp_sel(double t) {
  if (TAB) return exp_tab(t);
  return exp_fast(t);
}

// Kernel-specific pieces.  pk0 holds e_j = ln(alpha) - |x~_j|^2/2 for SE and |x~_j|^2 for Matern; hq = -|x~|^2/2.
// Returns the value weight kv (k(x, X_j)) and the gradient weight kb (d k / d x_d = kb * (x~_jd - x~_d) / l_d).
template <int KERNEL, bool TAB = false>
__device__ __forceinline__ void kernel_pair(double dot, double pk0, double hq, double alpha, double& kv, double& kb) {
  if (KERNEL == CMOE_KERNEL_SQUARE_EXPONENTIAL) {
    kv = exp_sel<TAB>(dot + (pk0 + hq));
    kb = kv;
  } else {
    const double r2 = fmax(0.0, pk0 - 2.0 * (hq + dot));
    const double ar = kSqrt5 * sqrt(r2);
    const double ee = alpha * exp_fast(-ar);
    kv = ee * (1.0 + ar + (5.0 / 3.0) * r2);
    kb = (5.0 / 3.0) * ee * (1.0 + ar);
  }
}

// 16-byte loads of the staged operands: LDS.128 when the pointer is known to be shared, LDG.128 (read-only path) else
template <bool SMEM>
__device__ __forceinline__ double2 ld2(const double* p) {
  if (SMEM) {
    return *reinterpret_cast<const double2*>(p);
  } else {
    return __ldg(reinterpret_cast<const double2*>(p));
  }
}

template <bool SMEM>
__device__ __forceinline__ double ld1(const double* p) {
  if (SMEM) return *p;
  return __ldg(p);
}

// x . y and beta - B . c over the staged operands (one dependent DFMA chain each; splitting them into even / odd
// half chains measured 4 % slower: more instructions, no latency win).
template <int DIM>
__device__ __forceinline__ double dot_dim(const double (&x)[DIM], const double (&y)[DIM], double init) {
  double dot = init;
#pragma unroll
  for (int d = 0; d < DIM; ++d) dot = fma(x[d], y[d], dot);
  return dot;
}

template <int QP, bool SMEM>
__device__ __forceinline__ double weight_row(const double* __restrict__ pk, double beta, const double (&c)[QP]) {
  double a = beta;
#pragma unroll
  for (int u = 0; u < QP; u += 2) {
    const double2 b = ld2<SMEM>(pk + 2 + u);
    a = fma(-b.x, c[u], a);
    a = fma(-b.y, c[u + 1], a);
  }
  return a;
}

// mu+(xq) - m  and  the scaled-gradient accumulators, for one query point (scaled coordinates xq).
template <int KERNEL, int DIM, int QP, bool SMEM>
__device__ __forceinline__ void eval_posterior(const double* __restrict__ Xt, const double* __restrict__ Pk,
                                               const double* __restrict__ Xu, int N, int U, double alpha,
                                               const double (&xq)[DIM], const double (&c)[QP], double& S0, double& SB,
                                               double (&s)[DIM]) {
  static_assert(DIM % 2 == 0 && QP % 2 == 0, "operands are moved as 16-byte pairs");
  double nq = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) nq = fma(xq[d], xq[d], nq);
  const double hq = -0.5 * nq;
  S0 = 0.0;
  SB = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) s[d] = 0.0;
#pragma unroll 2
  for (int j = 0; j < N; ++j) {
    const double* xj = Xt + j * DIM;
    const double* pk = Pk + j * (QP + 2);
    double xv[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d += 2) {
      const double2 v = ld2<SMEM>(xj + d);
      xv[d] = v.x;
      xv[d + 1] = v.y;
    }
    const double2 h = ld2<SMEM>(pk);  // (e_j, beta_j)
    const double dot = dot_dim<DIM>(xq, xv, 0.0);
    const double a = weight_row<QP, SMEM>(pk, h.y, c);
    double kv, kb;
    kernel_pair<KERNEL, SMEM>(dot, h.x, hq, alpha, kv, kb);
    S0 = fma(a, kv, S0);
    const double wb = a * kb;
    if (KERNEL != CMOE_KERNEL_SQUARE_EXPONENTIAL) SB += wb;
#pragma unroll
    for (int d = 0; d < DIM; ++d) s[d] = fma(wb, xv[d], s[d]);
  }
  for (int u = 0; u < U; ++u) {
    const double* xu = Xu + u * (DIM + 2);
    double xv[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d += 2) {
      const double2 v = ld2<SMEM>(xu + d);
      xv[d] = v.x;
      xv[d + 1] = v.y;
    }
    const double2 h = ld2<SMEM>(xu + DIM);
    double dot = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) dot = fma(xq[d], xv[d], dot);
    double kv, kb;
    kernel_pair<KERNEL, SMEM>(dot, h.x, hq, alpha, kv, kb);
    double cu = 0.0;
#pragma unroll
    for (int v = 0; v < QP; ++v)
      if (v == u) cu = c[v];
    S0 = fma(cu, kv, S0);
    const double wb = cu * kb;
    if (KERNEL != CMOE_KERNEL_SQUARE_EXPONENTIAL) SB += wb;
#pragma unroll
    for (int d = 0; d < DIM; ++d) s[d] = fma(wb, xv[d], s[d]);
  }
  if (KERNEL == CMOE_KERNEL_SQUARE_EXPONENTIAL) SB = S0;
  if (SMEM && nq > kFarRadius * kFarRadius) {
    // farther than kFarRadius length scales from everything: every kernel value is exactly 0 (see kFastPathRadius)
    S0 = 0.0;
    SB = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) s[d] = 0.0;
  }
}

// Batched backtracking for the SquareExponential kernel.  The reference's line search tries x + a_k g for
// a_k = a_0 2^-k one value-evaluation at a time (gpp_optimization.hpp:750-765).  Along a line the SE kernel factorises:
//     ln k(x + a g, X_j) = ln k(x, X_j) + a p_j - a^2 |g~|^2 / 2 ,   p_j = g~ . (X~_j - x~)
// so with F_j = exp(a_min p_j) the kernel at every trial point is k(x, X_j) F_j^(2^m) times a j-independent factor:
// two exps per training point give all KB trial values by repeated squaring (1 mul + 1 fma per trial) instead of one
// full evaluation (dot + weights + exp) per trial.  S[k] = sum_j a_j k(x, X_j) exp(a_k p_j) for a_k = a_min 2^(KB-1-k'),
// returned in trial order (S[0] <-> largest step); pmax_hi (high word of max_j |a_min p_j|) lets the caller reject
// batches whose factors could leave the double range (it then falls back to one-at-a-time evaluations).
constexpr int kLineBatch = 8;  // trial step sizes per batch (6 measured +1.5 %, but a second batch costs a whole pass)

template <int DIM, int QP, bool SMEM>
__device__ __forceinline__ void eval_line(const double* __restrict__ Xt, const double* __restrict__ Pk,
                                          const double* __restrict__ Xu, int N, int U, const double (&xb)[DIM],
                                          const double (&gt)[DIM], const double (&c)[QP], double alpha_min,
                                          double (&S)[kLineBatch], int& pmax_hi) {
  double nq = 0.0, xg = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    nq = fma(xb[d], xb[d], nq);
    xg = fma(xb[d], gt[d], xg);
  }
  const double hq = -0.5 * nq;
  double ga[DIM];  // a_min g~ : p_j is only ever used scaled by a_min
#pragma unroll
  for (int d = 0; d < DIM; ++d) ga[d] = alpha_min * gt[d];
  const double xga = -alpha_min * xg;
#pragma unroll
  for (int k = 0; k < kLineBatch; ++k) S[k] = 0.0;
  pmax_hi = 0;  // max over rows of the high word of |a_min p_j| (integer ALU; monotone in |p|, NaN/inf sort last)
#pragma unroll 2
  for (int j = 0; j < N; ++j) {
    const double* xj = Xt + j * DIM;
    const double* pk = Pk + j * (QP + 2);
    double xv[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d += 2) {
      const double2 v = ld2<SMEM>(xj + d);
      xv[d] = v.x;
      xv[d + 1] = v.y;
    }
    const double2 h = ld2<SMEM>(pk);  // (e_j, beta_j)
    const double dot = dot_dim<DIM>(xb, xv, 0.0);
    const double pj = dot_dim<DIM>(ga, xv, xga);
    const double a = weight_row<QP, SMEM>(pk, h.y, c);
    pmax_hi = max(pmax_hi, __double2hiint(pj) & 0x7fffffff);
    const double w = a * exp_sel<SMEM>(dot + (h.x + hq));
    double G = exp_sel<SMEM>(pj);
#pragma unroll
    for (int k = kLineBatch - 1; k >= 0; --k) {
      S[k] = fma(w, G, S[k]);
      if (k > 0) G *= G;
    }
  }
  for (int u = 0; u < U; ++u) {
    const double* xu = Xu + u * (DIM + 2);
    double xv[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d += 2) {
      const double2 v = ld2<SMEM>(xu + d);
      xv[d] = v.x;
      xv[d + 1] = v.y;
    }
    const double2 h = ld2<SMEM>(xu + DIM);
    double dot = 0.0, pj = xga;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      dot = fma(xb[d], xv[d], dot);
      pj = fma(ga[d], xv[d], pj);
    }
    double cu = 0.0;
#pragma unroll
    for (int v = 0; v < QP; ++v)
      if (v == u) cu = c[v];
    pmax_hi = max(pmax_hi, __double2hiint(pj) & 0x7fffffff);
    const double w = cu * exp_sel<SMEM>(dot + (h.x + hq));
    double G = exp_sel<SMEM>(pj);
#pragma unroll
    for (int k = kLineBatch - 1; k >= 0; --k) {
      S[k] = fma(w, G, S[k]);
      if (k > 0) G *= G;
    }
  }
}

// Batched backtracking for the Matern-5/2 kernel.  No factorisation along the line here, but the trial points
// x + a_k g share everything that depends on the training point only: with r0^2 = |x~ - X~_j|^2 and
// p_j = g~ . (X~_j - x~),   |x~ + a g~ - X~_j|^2 = r0^2 - 2 a p_j + a^2 |g~|^2 ,
// so the operand loads, the two dot products and the Q weight FMAs are paid once per step and each trial costs one
// sqrt, one exp and a handful of FMAs, value only (a one-at-a-time trial also accumulates the DIM gradient sums it
// almost never uses).  S[k] <-> a_k = a0 2^-k, k < KB.
#ifndef CMOE_LINE_BATCH_M
#define CMOE_LINE_BATCH_M 6
#endif
constexpr int kLineBatchM = CMOE_LINE_BATCH_M;
static_assert(kLineBatchM <= kLineBatch, "S is sized for the SquareExponential batch");

// sqrt(p) for a normal p > 0: MUFU.RSQ64H seed + two coupled Goldschmidt steps (seven FP64 operations, no special-case
// branch; within ~1 ulp)
__device__ __forceinline__ double sqrt_seeded(double p) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(p));
  double g = p * y0, h = 0.5 * y0;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  return fma(g, r, g);
}

template <int DIM, int QP, bool SMEM>
__device__ __forceinline__ void eval_line_matern(const double* __restrict__ Xt, const double* __restrict__ Pk,
                                                 const double* __restrict__ Xu, int N, int U, double alpha,
                                                 const double (&xb)[DIM], const double (&gt)[DIM],
                                                 const double (&c)[QP], double a0, double (&S)[kLineBatch]) {
  constexpr int KB = kLineBatchM;
  double nq = 0.0, xg = 0.0, gg = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    nq = fma(xb[d], xb[d], nq);
    xg = fma(xb[d], gt[d], xg);
    gg = fma(gt[d], gt[d], gg);
  }
  const double hq = -0.5 * nq;
  double m2a[KB], aag[KB];
  {
    double ak = a0;
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      m2a[k] = -2.0 * ak;
      aag[k] = ak * ak * gg;
      ak *= 0.5;
      S[k] = 0.0;
    }
  }
  auto trials = [&](double r2b, double pj, double w) {
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      // clamped away from 0 for the seed (Matern-5/2 is flat at r = 0: the value is unchanged to the last bit) and,
      // on the table-exp path, from above so that the argument stays inside the table exp's range contract
      const double r2 = fmax(1.0e-280, fma(m2a[k], pj, r2b) + aag[k]);
      double ar = kSqrt5 * sqrt_seeded(r2);
      if (SMEM) ar = fmin(ar, 1.0e6);
      const double we = w * exp_sel<SMEM>(-ar);
      S[k] = fma(we, fma(5.0 / 3.0, r2, 1.0 + ar), S[k]);
    }
  };
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    const double* xj = Xt + j * DIM;
    const double* pk = Pk + j * (QP + 2);
    double xv[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d += 2) {
      const double2 v = ld2<SMEM>(xj + d);
      xv[d] = v.x;
      xv[d + 1] = v.y;
    }
    const double2 h = ld2<SMEM>(pk);  // (|X~_j|^2, beta_j)
    const double dot = dot_dim<DIM>(xb, xv, 0.0);
    const double pj = dot_dim<DIM>(gt, xv, -xg);
    const double a = weight_row<QP, SMEM>(pk, h.y, c);
    trials(fmax(0.0, h.x - 2.0 * (hq + dot)), pj, a * alpha);
  }
  for (int u = 0; u < U; ++u) {
    const double* xu = Xu + u * (DIM + 2);
    double xv[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d += 2) {
      const double2 v = ld2<SMEM>(xu + d);
      xv[d] = v.x;
      xv[d + 1] = v.y;
    }
    const double2 h = ld2<SMEM>(xu + DIM);
    const double dot = dot_dim<DIM>(xb, xv, 0.0);
    const double pj = dot_dim<DIM>(gt, xv, -xg);
    double cu = 0.0;
#pragma unroll
    for (int v = 0; v < QP; ++v)
      if (v == u) cu = c[v];
    trials(fmax(0.0, h.x - 2.0 * (hq + dot)), pj, cu * alpha);
  }
}

// Kernel pieces for the general path: kv = k(x, X_j) (value row), kb = factor of the first-derivative rows,
// kc = factor of d kb / d x  (SE: all three equal k; Matern-5/2: cov0, first_derivative_part, alpha_exp_part of
// gpp_covariance.cpp:353-355).
template <int KERNEL>
__device__ __forceinline__ void kernel_triple(double dot, double pk0, double hq, double alpha, double& kv, double& kb,
                                              double& kc) {
  if (KERNEL == CMOE_KERNEL_SQUARE_EXPONENTIAL) {
    kv = exp_fast(dot + (pk0 + hq));
    kb = kv;
    kc = kv;
  } else {
    const double r2 = fmax(0.0, pk0 - 2.0 * (hq + dot));
    const double ar = kSqrt5 * sqrt(r2);
    const double ee = alpha * exp_fast(-ar);
    kv = ee * (1.0 + ar + (5.0 / 3.0) * r2);
    kb = (5.0 / 3.0) * ee * (1.0 + ar);
    kc = (25.0 / 3.0) * ee;
  }
}

constexpr int kMaxG = 8;
// points per trip of the general path's row loops: the loop body is one long dependent chain per point (weight loads
// from L2/HBM, dot, exp), so independent points are interleaved to cover the latencies
#ifndef CMOE_GEN_UNROLL
#define CMOE_GEN_UNROLL 4
#endif
constexpr int kGenUnroll = CMOE_GEN_UNROLL;

// Weight stream of the general path.  Each lane re-reads its weight column (N(1+g) doubles, global memory) on every
// evaluation; a plain load loop leaves the warp waiting on L2/HBM for most of its time (ncu: 74 % of the issue
// slots stalled on the long scoreboard with 8 warps per SM).  The column is therefore pulled through a per-lane ring in
// shared memory with LDGSTS (cp.async, 8 bytes per lane = 256 bytes per warp and row), kRingDepth stages of `P` points
// deep, and the arithmetic reads the ring.  Every lane copies and reads only its own slots, so cp.async.wait_group is
// the only synchronisation.  body(j, w): w[m * kMcThreads] is weight m of point j.
constexpr int kRingRows = 80;   // ring slots per lane (rows of the weight column)
constexpr int kRingDepth = 8;   // stages in flight
template <typename F>
__device__ __forceinline__ void stream_point_weights(const double* aw, double* ring, int N, int b1, F&& body) {
  constexpr int T = kMcThreads;
  const int P = max(1, kRingRows / (kRingDepth * b1));  // points per stage
  const int SR = P * b1, R = N * b1;
  const int nst = (N + P - 1) / P;
  auto issue = [&](int st) {
    if (st < nst) {
      const int r0 = st * SR;
      const int cnt = min(SR, R - r0);
      double* dst = ring + (st & (kRingDepth - 1)) * SR * T;
      const double* src = aw + static_cast<size_t>(r0) * T;
      for (int k = 0; k < cnt; ++k) cp_async8(dst + k * T, src + static_cast<size_t>(k) * T, true);
    }
    cp_async_commit();
  };
  for (int st = 0; st < kRingDepth - 1; ++st) issue(st);
  for (int st = 0; st < nst; ++st) {
    issue(st + kRingDepth - 1);  // refills the slot consumed in the previous trip
    cp_async_wait<kRingDepth - 1>();
    const double* w = ring + (st & (kRingDepth - 1)) * SR * T;
    const int j1 = min(N, (st + 1) * P);
#pragma unroll 2
    for (int j = st * P; j < j1; ++j, w += b1 * T) body(j, w);
  }
  cp_async_wait<0>();
}

// General evaluation with derivative observations (d-KG): every training point owns 1+g rows
//   K((j,0), x) = kv ,  K((j,m), x) = kb (x~_t - X~_jt) / l_t   (t = derivs[m-1]; the 1/l_t is folded into the pack)
//   mu+(x) - m = sum_j [ a_j0 kv + kb sum_m a~_jm (x~_t - X~_jt) ]  with  a_(j,m) = beta_(j,m) - B_(j,m),: . c
//   d/dx~_d   = (X~_jd - x~_d) (a_j0 kb + kc wsum) + kb a~_jm [d == t]
// The weights a_(j,m) come from this lane's ring slots (aj), the point operands from shared memory when they were
// parked there (XS) and through the read-only path otherwise.
template <int KERNEL, int DIM, int QP, bool XS>
__device__ __forceinline__ void eval_posterior_gen(const KgMcParams& prm, const double* __restrict__ Xt,
                                                   const double* __restrict__ Pk, const double* __restrict__ Xu,
                                                   const double (&xq)[DIM], const double* aw, double* ring,
                                                   const double* __restrict__ cl, double& S0, double& SB,
                                                   double (&s)[DIM], double (&em)[kMaxG]) {
  const int N = prm.N, U = prm.U, g = prm.g, stride = prm.pk_stride;
  double nq = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) nq = fma(xq[d], xq[d], nq);
  const double hq = -0.5 * nq;
  // query coordinates at the derivative indices (compile-time register indexing only)
  double xm[kMaxG];
#pragma unroll
  for (int m = 0; m < kMaxG; ++m) {
    xm[m] = 0.0;
    em[m] = 0.0;
    if (m < g) {
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        if (d == prm.derivs[m]) xm[m] = xq[d];
    }
  }
  S0 = 0.0;
  SB = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) s[d] = 0.0;
  stream_point_weights(aw, ring, N, 1 + g, [&](int j, const double* aj) {
    const double* xj = Xt + static_cast<size_t>(j) * DIM;
    const double* pk = Pk + static_cast<size_t>(j) * stride;
    double xv[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d += 2) {
      const double2 v = ld2<XS>(xj + d);
      xv[d] = v.x;
      xv[d + 1] = v.y;
    }
    double dot = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) dot = fma(xq[d], xv[d], dot);
    double kv, kb, kc;
    kernel_triple<KERNEL>(dot, ld1<XS>(pk), hq, prm.alpha, kv, kb, kc);
    const double a0 = aj[0];  // this lane's weights of point j (ring slots)
    double wsum = 0.0;
#pragma unroll
    for (int m = 0; m < kMaxG; ++m) {
      if (m < g) {
        const double am = aj[(m + 1) * kMcThreads];
        wsum = fma(am, xm[m] - ld1<XS>(xj + prm.derivs[m]), wsum);
        em[m] = fma(kb, am, em[m]);
      }
    }
    S0 = fma(a0, kv, S0);
    S0 = fma(kb, wsum, S0);
    const double wb = fma(kc, wsum, a0 * kb);
    SB += wb;
#pragma unroll
    for (int d = 0; d < DIM; ++d) s[d] = fma(wb, xv[d], s[d]);
  });
  const int bs = 1 + g;
  for (int u = 0; u < U; ++u) {
    const double* xu = Xu + static_cast<size_t>(u) * (DIM + 2);
    double xv[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d += 2) {
      const double2 v = ld2<false>(xu + d);
      xv[d] = v.x;
      xv[d + 1] = v.y;
    }
    double dot = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) dot = fma(xq[d], xv[d], dot);
    double kv, kb, kc;
    kernel_triple<KERNEL>(dot, __ldg(xu + DIM), hq, prm.alpha, kv, kb, kc);
    const double a0 = cl[u * bs];
    double wsum = 0.0;
#pragma unroll
    for (int m = 0; m < kMaxG; ++m) {
      if (m < g) {
        const double am = cl[u * bs + 1 + m] * prm.inv_len[prm.derivs[m]];
        wsum = fma(am, xm[m] - __ldg(xu + prm.derivs[m]), wsum);
        em[m] = fma(kb, am, em[m]);
      }
    }
    S0 = fma(a0, kv, S0);
    S0 = fma(kb, wsum, S0);
    const double wb = fma(kc, wsum, a0 * kb);
    SB += wb;
#pragma unroll
    for (int d = 0; d < DIM; ++d) s[d] = fma(wb, xv[d], s[d]);
  }
}

// Batched backtracking with derivative observations (SquareExponential): along x + a g the value rows scale as in
// eval_line and the derivative rows' factor (x~_t - X~_jt) is linear in a, so with
//     u_j = sum_m a~_jm (x~_t - X~_jt),  v_j = sum_m a~_jm g~_t,  k0_j = k(x, X_j),  G_jk = exp(a_k p_j)
// the trial values are  E_k [ sum_j k0_j G_jk (a_j0 + u_j) + a_k sum_j k0_j G_jk v_j ]:  S[k] and T[k] below.
// The (1+g) Q weight FMAs per point — the dominant cost of this path — are paid once per step instead of once per trial.
template <int DIM, int QP, bool XS>
__device__ __forceinline__ void eval_line_gen(const KgMcParams& prm, const double* __restrict__ Xt,
                                              const double* __restrict__ Pk, const double* __restrict__ Xu,
                                              const double (&xb)[DIM], const double (&gt)[DIM], const double* aw,
                                              double* ring, const double* __restrict__ cl, double alpha_min,
                                              double (&S)[kLineBatch], double (&T)[kLineBatch], int& pmax_hi) {
  const int N = prm.N, U = prm.U, g = prm.g, stride = prm.pk_stride;
  double nq = 0.0, xg = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    nq = fma(xb[d], xb[d], nq);
    xg = fma(xb[d], gt[d], xg);
  }
  const double hq = -0.5 * nq;
  double ga[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) ga[d] = alpha_min * gt[d];
  const double xga = -alpha_min * xg;
  // base coordinates and direction at the derivative indices (compile-time register indexing only)
  double xm[kMaxG], gm[kMaxG];
#pragma unroll
  for (int m = 0; m < kMaxG; ++m) {
    xm[m] = 0.0;
    gm[m] = 0.0;
    if (m < g) {
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        if (d == prm.derivs[m]) {
          xm[m] = xb[d];
          gm[m] = gt[d];
        }
    }
  }
#pragma unroll
  for (int k = 0; k < kLineBatch; ++k) S[k] = T[k] = 0.0;
  pmax_hi = 0;
  stream_point_weights(aw, ring, N, 1 + g, [&](int j, const double* aj) {
    const double* xj = Xt + static_cast<size_t>(j) * DIM;
    const double* pk = Pk + static_cast<size_t>(j) * stride;
    double xv[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d += 2) {
      const double2 v = ld2<XS>(xj + d);
      xv[d] = v.x;
      xv[d + 1] = v.y;
    }
    double dot = 0.0, pj = xga;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      dot = fma(xb[d], xv[d], dot);
      pj = fma(ga[d], xv[d], pj);
    }
    const double a0 = aj[0];
    double us = 0.0, vs = 0.0;
#pragma unroll
    for (int m = 0; m < kMaxG; ++m) {
      if (m < g) {
        const double am = aj[(m + 1) * kMcThreads];
        us = fma(am, xm[m] - ld1<XS>(xj + prm.derivs[m]), us);
        vs = fma(am, gm[m], vs);
      }
    }
    pmax_hi = max(pmax_hi, __double2hiint(pj) & 0x7fffffff);
    const double k0 = exp_fast(dot + (ld1<XS>(pk) + hq));
    const double w = k0 * (a0 + us), wv = k0 * vs;
    double G = exp_fast(pj);
#pragma unroll
    for (int k = kLineBatch - 1; k >= 0; --k) {
      S[k] = fma(w, G, S[k]);
      T[k] = fma(wv, G, T[k]);
      if (k > 0) G *= G;
    }
  });
  const int bs = 1 + g;
  for (int u = 0; u < U; ++u) {
    const double* xu = Xu + static_cast<size_t>(u) * (DIM + 2);
    double xv[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d += 2) {
      const double2 v = ld2<false>(xu + d);
      xv[d] = v.x;
      xv[d + 1] = v.y;
    }
    double dot = 0.0, pj = xga;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      dot = fma(xb[d], xv[d], dot);
      pj = fma(ga[d], xv[d], pj);
    }
    const double a0 = cl[u * bs];
    double us = 0.0, vs = 0.0;
#pragma unroll
    for (int m = 0; m < kMaxG; ++m) {
      if (m < g) {
        const double am = cl[u * bs + 1 + m] * prm.inv_len[prm.derivs[m]];
        us = fma(am, xm[m] - __ldg(xu + prm.derivs[m]), us);
        vs = fma(am, gm[m], vs);
      }
    }
    pmax_hi = max(pmax_hi, __double2hiint(pj) & 0x7fffffff);
    const double k0 = exp_fast(dot + (__ldg(xu + DIM) + hq));
    const double w = k0 * (a0 + us), wv = k0 * vs;
    double G = exp_fast(pj);
#pragma unroll
    for (int k = kLineBatch - 1; k >= 0; --k) {
      S[k] = fma(w, G, S[k]);
      T[k] = fma(wv, G, T[k]);
      if (k > 0) G *= G;
    }
  }
}

// TensorProductDomain::LimitUpdate, gpp_domain.cpp:64-104, for one coordinate
__device__ __forceinline__ double limit_step(double step, double x, double lo, double hi, double mrc) {
  double dist = fmin(x - lo, hi - x);
  if (fabs(step) > mrc * dist) step = copysign(mrc * dist, step);
  const double next = x + step;
  if (next < lo || next > hi) {
    if (next < lo) {
      dist = lo - x;
      step = (x + step * 0.5 < lo) ? dist * 0.5 : step * 0.5;
    } else {
      dist = hi - x;
      step = (x + step * 0.5 > hi) ? dist * 0.5 : step * 0.5;
    }
  }
  return step;
}

// General path: the weights a = beta~ - B~ c depend on the sample only, not on the query point, so they are formed once
// when a lane takes a new sample and parked in that lane's column of the CTA's scratch (global memory, streamed back
// with L2-only loads by every evaluation of the sample: (1+g) loads per training point instead of (1+g)(Q+1) loads
// and (1+g) Q FMAs).  The warp fills the column of each requesting lane together, rows strided over the lanes, so the
// table reads are coalesced; the FMA order per row (beta first, then u ascending) is the one the per-evaluation code had.
template <int QP>
__device__ __forceinline__ void fill_sample_weights(const double* __restrict__ Wt, int R,
                                                    const double* __restrict__ recC, double* aw_warp, unsigned want,
                                                    int sample, int lane) {
  while (want) {
    const int b = __ffs(want) - 1;
    want &= want - 1;
    const int sb = __shfl_sync(0xffffffffu, sample, b);
    double cb[QP];
#pragma unroll
    for (int u = 0; u < QP; ++u) cb[u] = __ldg(recC + static_cast<size_t>(sb) * QP + u);
    double* col = aw_warp + b;
#pragma unroll 4
    for (int r = lane; r < R; r += 32) {
      double a = __ldg(Wt + r);
#pragma unroll
      for (int u = 0; u < QP; ++u) a = fma(-__ldg(Wt + static_cast<size_t>(1 + u) * R + r), cb[u], a);
      col[static_cast<size_t>(r) * kMcThreads] = a;
    }
  }
  __syncwarp();
}

// The same column formed by its own lane (table loads are warp-uniform).  Costs R Q FMAs per lane however many lanes
// take part, so it is used when at least kOwnFillLanes do; a lone lane is served by the whole warp instead.
constexpr int kOwnFillLanes = 6;
template <int QP>
__device__ __forceinline__ void fill_own_weights(const double* __restrict__ Wt, int R, const double (&c)[QP],
                                                 double* aw) {
#pragma unroll 4
  for (int r = 0; r < R; ++r) {
    double a = __ldg(Wt + r);
#pragma unroll
    for (int u = 0; u < QP; ++u) a = fma(-__ldg(Wt + static_cast<size_t>(1 + u) * R + r), c[u], a);
    aw[static_cast<size_t>(r) * kMcThreads] = a;
  }
}

enum : int { ST_FETCH = 0, ST_INIT = 1, ST_TRIAL = 2, ST_LIMIT = 3, ST_DONE = 4, ST_LINE = 5 };


// The per-lane line-search state machine.  Live across evaluations: c, the base point xb with f and grad f there,
// the step size and a few counters; the start point of the current restart run is parked in the sample's output slot.
//
// A round of the warp is either a POINT round (mu+ and its gradient at one query point per lane: INIT / TRIAL / LIMIT
// lanes) or, for the SquareExponential fast path, a LINE round (all backtracking trials of a step at once, eval_line).
// Which one runs is a warp vote; lanes waiting for the other kind idle for that round and are in phase again after it
// (a sample alternates LINE, POINT, LINE, ... so the vote keeps a warp in lock-step once it is).
template <int KERNEL, int DIM, int QP, bool SMEM, bool GEN>
__device__ __forceinline__ void kg_mc_body(const KgMcParams& prm, const double* __restrict__ Xt,
                                           const double* __restrict__ Pk, const double* __restrict__ Xu, int cand,
                                           int s_begin, int s_end, int* next_sample, double* aw = nullptr,
                                           const double* __restrict__ Wt = nullptr, double* ring = nullptr) {
  constexpr bool SE = (KERNEL == CMOE_KERNEL_SQUARE_EXPONENTIAL);
  constexpr bool LINE = SE || !GEN;  // Matern-5/2 batches its trials too (eval_line_matern) unless derivative rows exist
  constexpr int KB = SE ? kLineBatch : kLineBatchM;  // trial step sizes per LINE round
  constexpr int ST_SEARCH = LINE ? ST_LINE : ST_TRIAL;  // how a step's backtracking starts
  const int N = prm.N, U = prm.U;
  const double* A = prm.A + static_cast<size_t>(cand) * prm.M * DIM;
  const double* recC = prm.recC + static_cast<size_t>(cand) * prm.num_mc * QP;
  const int* recStart = prm.recStart + static_cast<size_t>(cand) * prm.num_mc;
  double* outVal = prm.outVal + static_cast<size_t>(cand) * prm.num_mc;
  double* outX = prm.outX + static_cast<size_t>(cand) * prm.num_mc * DIM;
  double* outH = prm.outH + static_cast<size_t>(cand) * prm.num_mc;

  int state = ST_FETCH;
  int sample = s_begin + threadIdx.x;
  double c[QP], xb[DIM], gb[DIM];
  double cl[GEN ? QP : 1];  // general path: run-time indexable copy of c for the union rows (lives in local memory)
  double fb = 0.0, alpha_n = 0.0, gnorm = 0.0;
  int step_i = 0, restart_i = 0, search = 0;
  unsigned n_evals = 0, n_steps = 0, n_point = 0, n_line = 0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) xb[d] = gb[d] = 0.0;
#pragma unroll
  for (int u = 0; u < QP; ++u) c[u] = 0.0;

  while (true) {
    bool took_sample = false;
    if (state == ST_FETCH) {
      if (sample < s_end) {
        if (prm.max_restarts > 0) {
          took_sample = true;
#pragma unroll
          for (int u = 0; u < QP; ++u) c[u] = recC[static_cast<size_t>(sample) * QP + u];
          if (GEN) {
#pragma unroll
            for (int u = 0; u < QP; ++u) cl[GEN ? u : 0] = c[u];
          }
          const double* a0 = A + static_cast<size_t>(recStart[sample]) * DIM;
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            xb[d] = a0[d];
            outX[static_cast<size_t>(sample) * DIM + d] = a0[d];  // start of restart run 0
          }
          step_i = 0;
          restart_i = 0;
          state = ST_INIT;
        } else {
          // ComputeOptimalPosteriorMean returns without touching its outputs (...optimization.cpp:424-426):
          // best_function_value stays 0 and the best point keeps its fill value 1.0
          outVal[sample] = 0.0;
          double nx = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            outX[static_cast<size_t>(sample) * DIM + d] = 1.0 * prm.inv_len[d];
            nx = fma(prm.inv_len[d], prm.inv_len[d], nx);
          }
          outH[sample] = -0.5 * nx;
          sample = atomicAdd(next_sample, 1);
        }
      } else {
        state = ST_DONE;
      }
    }
    if (__all_sync(0xffffffffu, state == ST_DONE)) break;
    if (GEN) {
      const unsigned want = __ballot_sync(0xffffffffu, took_sample);
      if (__popc(want) >= kOwnFillLanes) {
        // many lanes at once (samples with equal round counts keep a warp's lanes in cohorts): every requesting lane
        // forms its own column from warp-uniform (broadcast) table loads; the column stores are full lines
        if (took_sample) fill_own_weights<QP>(Wt, prm.N * (1 + prm.g), c, aw);
        __syncwarp();
      } else if (want) {
        fill_sample_weights<QP>(Wt, prm.N * (1 + prm.g), recC, aw - (threadIdx.x & 31), want, sample,
                                threadIdx.x & 31);
      }
    }

    bool line_round = false;
    if (LINE) {
      const unsigned wl = __ballot_sync(0xffffffffu, state == ST_LINE);
      const unsigned wp = __ballot_sync(0xffffffffu, state == ST_INIT || state == ST_TRIAL || state == ST_LIMIT);
      line_round = __popc(wl) > __popc(wp);
    }
    bool finish_run = false;

    if (LINE && line_round) {
      // ---- all backtracking trials of this step in one pass (warp-uniform) ----
      double xt[DIM], gt[DIM], S[kLineBatch];
      int pmax_hi;
      double gg = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        xt[d] = xb[d] * prm.inv_len[d];
        gt[d] = gb[d] * prm.inv_len[d];
        gg = fma(gt[d], gt[d], gg);
      }
      constexpr double kTop = static_cast<double>(1 << (KB - 1));
      double T[GEN ? kLineBatch : 1];
      if (!SE) {
        pmax_hi = 0;  // no factor can leave the double range on this path
        eval_line_matern<DIM, QP, SMEM>(Xt, Pk, Xu, N, U, prm.alpha, xt, gt, c, alpha_n, S);
      } else if (GEN) {
        eval_line_gen<DIM, QP, SMEM>(prm, Xt, Pk, Xu, xt, gt, aw, ring, cl, alpha_n * (1.0 / kTop), S,
                               reinterpret_cast<double (&)[kLineBatch]>(T), pmax_hi);
      } else {
        eval_line<DIM, QP, SMEM>(Xt, Pk, Xu, N, U, xt, gt, c, alpha_n * (1.0 / kTop), S, pmax_hi);
      }
      if (state == ST_LINE) {
        n_line += 1;
        // safe iff a_0 |p_j| = 2^(KB-1) |a_min p_j| < 512 for every row: every factor exp(a_k p_j) stays in range
        double nb2 = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) nb2 = fma(xt[d], xt[d], nb2);
        if (pmax_hi >= __double2hiint(512.0 / kTop) || !(nb2 <= kFarRadius * kFarRadius)) {
          state = ST_TRIAL;  // factors could leave the double range: this step backtracks one evaluation at a time
        } else {
          int kacc = -1;
          double ak = alpha_n, a_acc = alpha_n;
#pragma unroll
          for (int k = 0; k < KB; ++k) {
            const double sk = GEN ? fma(ak, T[GEN ? k : 0], S[k]) : S[k];
            const double fq = SE ? -(prm.mean + exp_fast(-0.5 * ak * ak * gg) * sk) : -(prm.mean + sk);
            // Armijo-type test of the reference: f(x + a g) - f(x) > 0.5 a |g|^2   (gpp_optimization.hpp:758)
            const bool ok = (search + k < 30) && ((fq - fb) > 0.5 * ak * gnorm);
            if (kacc < 0 && ok) {
              kacc = k;
              a_acc = ak;
            }
            ak *= 0.5;
          }
          if (kacc >= 0) {
            n_evals += kacc + 1;
            search += kacc;
            alpha_n = a_acc;
            state = ST_LIMIT;  // value and gradient at the domain-limited point decide the step (:767-781)
          } else {
            n_evals += min(KB, 30 - search);
            search += KB;
            alpha_n *= 1.0 / (2.0 * kTop);
            if (search >= 30) finish_run = true;  // exhausted: reject and stop this run (:778-781)
          }
        }
      }
    } else {
      // ---- query point of this lane: base (INIT), base + alpha*grad (TRIAL), base + limited step (LIMIT) ----
      double xq[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        double st = 0.0;
        if (state == ST_TRIAL || state == ST_LIMIT) st = alpha_n * gb[d];
        if (state == ST_LIMIT) st = (d < prm.ps) ? limit_step(st, xb[d], prm.lo[d], prm.hi[d], prm.mrc) : 0.0;
        xq[d] = (xb[d] + st) * prm.inv_len[d];
      }

      // ---- the expensive, warp-uniform part ----
      double S0, SB, s[DIM];
      double em[kMaxG];
      if (GEN) {
        eval_posterior_gen<KERNEL, DIM, QP, SMEM>(prm, Xt, Pk, Xu, xq, aw, ring, cl, S0, SB, s, em);
        // kb a~_jm lands on coordinate derivs[m]: fold it into s so that the common gradient formula below holds
        // (grad_d = inv_len_d (s_d - x~_d SB)), using compile-time register indices only
#pragma unroll
        for (int m = 0; m < kMaxG; ++m)
          if (m < prm.g) {
#pragma unroll
            for (int d = 0; d < DIM; ++d)
              if (d == prm.derivs[m]) s[d] += em[m];
          }
      } else {
        eval_posterior<KERNEL, DIM, QP, SMEM>(Xt, Pk, Xu, N, U, prm.alpha, xq, c, S0, SB, s);
      }
      if (state == ST_INIT || state == ST_TRIAL || state == ST_LIMIT) {
        n_evals += 1;
        n_point += 1;
      }
      const double fq = -(prm.mean + S0);

      // ---- per-lane transitions (cheap) ----
      bool accept = false;
      if (state == ST_INIT) {
        fb = fq;
#pragma unroll
        for (int d = 0; d < DIM; ++d) gb[d] = (d < prm.ps) ? -prm.inv_len[d] * (s[d] - xq[d] * SB) : 0.0;
        if (prm.max_steps <= 0) {
          finish_run = true;
        } else {
          step_i = -1;  // becomes 0 in the common "start a step" block below
          accept = true;
        }
      } else if (state == ST_TRIAL) {
        // Armijo-type test of the reference: f(x + a g) - f(x) > 0.5 a |g|^2   (gpp_optimization.hpp:758)
        const bool ok = (fq - fb) > 0.5 * alpha_n * gnorm;
        if (!ok) {
          alpha_n *= 0.5;
          search += 1;
          if (search >= 30) finish_run = true;  // exhausted: reject and stop this run (:778-781)
        } else {
          // limit the accepted step to the domain; if unchanged, this evaluation IS the final evaluation of the step
          bool same = true;
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            const double raw = alpha_n * gb[d];
            const double lim = (d < prm.ps) ? limit_step(raw, xb[d], prm.lo[d], prm.hi[d], prm.mrc) : 0.0;
            same = same && (lim == raw);
          }
          if (same) {
            if (fq <= fb) {
              finish_run = true;
            } else {
              accept = true;
            }
          } else {
            state = ST_LIMIT;  // one more evaluation at the limited point
          }
        }
      } else if (state == ST_LIMIT) {
        if (fq <= fb) {
          finish_run = true;  // no increase: restore the base point and stop (:778-781)
        } else {
          accept = true;
        }
      }
      if (accept) {
        // the evaluated query point becomes the base point; start the next step (or finish the run)
        double ns = 0.0;
        if (state != ST_INIT) {
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            double st = alpha_n * gb[d];
            if (state == ST_LIMIT) st = (d < prm.ps) ? limit_step(st, xb[d], prm.lo[d], prm.hi[d], prm.mrc) : 0.0;
            xb[d] += st;
            ns = fma(st, st, ns);
          }
#pragma unroll
          for (int d = 0; d < DIM; ++d) gb[d] = (d < prm.ps) ? -prm.inv_len[d] * (s[d] - xq[d] * SB) : 0.0;
          fb = fq;
          n_steps += 1;
        }
        step_i += 1;
        if (state != ST_INIT && (sqrt(ns) < prm.step_tol || step_i >= prm.max_steps)) {
          finish_run = true;
        } else {
          alpha_n = prm.alpha0[step_i];
          search = 0;
          gnorm = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) gnorm = fma(gb[d], gb[d], gnorm);
          state = ST_SEARCH;
        }
      }
    }
    if (finish_run) {
      // restart logic of GradientDescentOptimizerLineSearch::Optimize (gpp_optimization.hpp:1255-1273)
      double nd2 = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const double df = outX[static_cast<size_t>(sample) * DIM + d] - xb[d];
        nd2 = fma(df, df, nd2);
      }
      restart_i += 1;
      if (sqrt(nd2) <= prm.tol || restart_i >= prm.max_restarts || prm.max_steps <= 0) {
        outVal[sample] = fb;
        double nx = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          const double xs = xb[d] * prm.inv_len[d];
          outX[static_cast<size_t>(sample) * DIM + d] = xs;
          nx = fma(xs, xs, nx);
        }
        outH[sample] = -0.5 * nx;
        sample = atomicAdd(next_sample, 1);
        state = ST_FETCH;
      } else {
#pragma unroll
        for (int d = 0; d < DIM; ++d) outX[static_cast<size_t>(sample) * DIM + d] = xb[d];
        step_i = 0;
        alpha_n = prm.alpha0[0];
        search = 0;
        gnorm = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) gnorm = fma(gb[d], gb[d], gnorm);
        state = ST_SEARCH;
      }
    }
  }
  // integer counters: atomics keep the totals deterministic
  n_evals = __reduce_add_sync(0xffffffffu, n_evals);
  n_steps = __reduce_add_sync(0xffffffffu, n_steps);
  n_point = __reduce_add_sync(0xffffffffu, n_point);
  n_line = __reduce_add_sync(0xffffffffu, n_line);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(prm.stats + 0, static_cast<unsigned long long>(n_evals));
    atomicAdd(prm.stats + 1, static_cast<unsigned long long>(n_steps));
    atomicAdd(prm.stats + 2, static_cast<unsigned long long>(n_point));
    atomicAdd(prm.stats + 3, static_cast<unsigned long long>(n_line));
  }
}

template <int KERNEL, int DIM, int QP>
__global__ void __launch_bounds__(kMcThreads, CMOE_MC_MINBLOCKS) kg_mc_kernel(const __grid_constant__ KgMcParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t mbar;
  __shared__ int next_sample;
  const int cand = blockIdx.y;
  const int N = prm.N, U = prm.U;
  const int s_begin = blockIdx.x * prm.chunk;
  const int s_end = min(prm.num_mc, s_begin + prm.chunk);
  const double* gXt = prm.Xt;
  const double* gPk = prm.Pk + static_cast<size_t>(cand) * N * (QP + 2);
  const double* gXu = prm.Xu + static_cast<size_t>(cand) * U * (DIM + 2);
  if (threadIdx.x == 0) next_sample = s_begin + blockDim.x;
  exp_table_stage();
  if (prm.use_smem) {
    // stage the per-candidate operands with TMA bulk copies (UBLKCP) signalled through an mbarrier
    double* sXt = reinterpret_cast<double*>(smem_raw);
    double* sPk = sXt + static_cast<size_t>(N) * DIM;
    double* sXu = sPk + static_cast<size_t>(N) * (QP + 2);
    const uint32_t bX = static_cast<uint32_t>(N) * DIM * 8u, bP = static_cast<uint32_t>(N) * (QP + 2) * 8u,
                   bU = static_cast<uint32_t>(U) * (DIM + 2) * 8u;
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&mbar, bX + bP + bU);
      tma_bulk_g2s(sXt, gXt, bX, &mbar);
      tma_bulk_g2s(sPk, gPk, bP, &mbar);
      tma_bulk_g2s(sXu, gXu, bU, &mbar);
    }
    mbar_wait(&mbar, 0);
    kg_mc_body<KERNEL, DIM, QP, true, false>(prm, sXt, sPk, sXu, cand, s_begin, s_end, &next_sample);
  } else {
    __syncthreads();
    kg_mc_body<KERNEL, DIM, QP, false, false>(prm, gXt, gPk, gXu, cand, s_begin, s_end, &next_sample);
  }
}

// General-path kernel (derivative observations): same state machine; the per-point operands are read through the
// read-only path and the per-sample weights from the lane's scratch column (fill_sample_weights).  Persistent grid: one
// CTA per scratch slot, work items (candidate, sample chunk) handed out by an atomic counter.
#ifndef CMOE_MC_GEN_MINBLOCKS
#define CMOE_MC_GEN_MINBLOCKS 2
#endif
template <int KERNEL, int DIM, int QP>
__global__ void __launch_bounds__(kMcThreads, CMOE_MC_GEN_MINBLOCKS)
    kg_mc_gen_kernel(const __grid_constant__ KgMcParams prm) {
  extern __shared__ __align__(16) double gen_ring[];  // [kRingRows][kMcThreads]
  __shared__ int next_sample;
  __shared__ int work_item;
  const int R = prm.N * (1 + prm.g);
  double* aw = prm.aw + static_cast<size_t>(blockIdx.x) * R * kMcThreads + threadIdx.x;
  // The weight stream passes through L1 on its way into the ring (8-byte cp.async is .ca only) and evicts everything
  // else, so the candidate-independent operands of every point — scaled coordinates and e_j — are parked behind the ring
  // in shared memory when they fit (read as broadcast LDS); e_j is the same for every candidate (candidate 0's pack).
  double* sX = gen_ring + static_cast<size_t>(kRingRows + kRingDepth) * kMcThreads;
  double* sE = sX + static_cast<size_t>(prm.N) * DIM;
  if (prm.stage_ops) {
    for (int e = threadIdx.x; e < prm.N * DIM; e += blockDim.x) sX[e] = prm.Xt[e];
    for (int j = threadIdx.x; j < prm.N; j += blockDim.x) {
      sE[2 * j] = prm.Pk[static_cast<size_t>(j) * prm.pk_stride];
      sE[2 * j + 1] = 0.0;
    }
    __syncthreads();
  }
  while (true) {
    if (threadIdx.x == 0) work_item = atomicAdd(prm.work, 1);
    __syncthreads();
    const int w = work_item;
    if (w >= prm.work_total) break;
    const int cand = w / prm.chunks;
    const int s_begin = (w % prm.chunks) * prm.chunk;
    const int s_end = min(prm.num_mc, s_begin + prm.chunk);
    if (threadIdx.x == 0) next_sample = s_begin + blockDim.x;
    __syncthreads();
    const double* gPk = prm.Pk + static_cast<size_t>(cand) * prm.N * prm.pk_stride;
    const double* gXu = prm.Xu + static_cast<size_t>(cand) * prm.U * (DIM + 2);
    const double* gWt = prm.Wt + static_cast<size_t>(cand) * (1 + QP) * R;
    if (prm.stage_ops) {
      kg_mc_body<KERNEL, DIM, QP, true, true>(prm, sX, sE, gXu, cand, s_begin, s_end, &next_sample, aw, gWt,
                                              gen_ring + threadIdx.x);
    } else {
      kg_mc_body<KERNEL, DIM, QP, false, true>(prm, prm.Xt, gPk, gXu, cand, s_begin, s_end, &next_sample, aw, gWt,
                                               gen_ring + threadIdx.x);
    }
    __syncthreads();  // next_sample / work_item are rewritten by thread 0
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Phase 2: per-candidate accumulations over the samples' minimisers for the envelope-theorem gradient
//   R[a][row] = sum_i c_ia k(row, x*_i)   (rows = training points then union points; fixed sample order)
//   Gu[u][d]  = sum_i c_iu kb(Xu_u, x*_i) x~*_id ,  GkB[u] = sum_i c_iu kb(Xu_u, x*_i)
// grid (row blocks, candidates); one thread per row.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kAccTile = 64;  // samples per shared-memory tile of kg_acc_kernel

// One thread per row of [training points; union points]; the samples' records (minimiser, c, -|x*|^2/2) stream through
// shared memory in double-buffered tiles (LDGSTS, 16-byte chunks) and are read back as warp-broadcast LDS.128.
// TAB selects the table exp under the fast path's range contract (kFastPathRadius).
// GEN (derivative observations): the rows are (point, m), m = 0..g, training rows first; row (p, m > 0) is the
// derivative of k(p, x*) with respect to coordinate t = derivs[m-1] of p:  kb (x~*_t - p~_t) / l_t.  The union rows'
// own-point gradient sums (Gu, GkB) are only produced for g == 0 (kg_g1_kernel does that part otherwise).
template <int KERNEL, int DIM, int QP, bool TAB, bool GEN = false>
__device__ __forceinline__ void kg_acc_body(const KgAccParams& prm, double* __restrict__ sbuf) {
  const int cand = blockIdx.y;
  const int grow = blockIdx.x * blockDim.x + threadIdx.x;
  const int b1 = GEN ? 1 + prm.g : 1;
  const int N = prm.N, U = prm.U, nrows = (N + U) * b1;
  const bool active = grow < nrows;
  const bool is_u = grow >= N * b1;
  const int row = GEN ? (is_u ? N + (grow - N * b1) / b1 : grow / b1) : grow;  // index into [training ; union] points
  const int mrow = GEN ? (is_u ? (grow - N * b1) % b1 : grow % b1) : 0;
  int tcoord = 0;
  double tscale = 0.0;  // 1 / l_t of a derivative row
  if (GEN && mrow > 0) {
    tcoord = prm.derivs[mrow - 1];
    tscale = prm.inv_len[tcoord];
  }
  double xr[DIM];
  double pk0 = 0.0;
  if (active) {
    const double* src = is_u ? (prm.Xu + (static_cast<size_t>(cand) * U + (row - N)) * (DIM + 2))
                             : (prm.Xt + static_cast<size_t>(row) * DIM);
#pragma unroll
    for (int d = 0; d < DIM; ++d) xr[d] = src[d];
    if (is_u) {
      pk0 = src[DIM];
    } else {
      double nrm = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) nrm = fma(xr[d], xr[d], nrm);
      pk0 = (KERNEL == CMOE_KERNEL_SQUARE_EXPONENTIAL) ? (log(prm.alpha) - 0.5 * nrm) : nrm;
    }
  } else {
#pragma unroll
    for (int d = 0; d < DIM; ++d) xr[d] = 0.0;
  }
  double acc[QP];
#pragma unroll
  for (int a = 0; a < QP; ++a) acc[a] = 0.0;
  double gu[DIM], gkb = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) gu[d] = 0.0;
  const double* xs = prm.outX + static_cast<size_t>(cand) * prm.num_mc * DIM;
  const double* cs = prm.recC + static_cast<size_t>(cand) * prm.num_mc * QP;
  const double* hs = prm.outH + static_cast<size_t>(cand) * prm.num_mc;
  const int uu = is_u ? (row - N) : 0;
  double xrt = 0.0;  // own coordinate along the derivative direction
  if (GEN) {
#pragma unroll
    for (int d = 0; d < DIM; ++d)
      if (d == tcoord) xrt = xr[d];
  }
  constexpr int TS = kAccTile;
  constexpr int kBuf = TS * (DIM + QP + 1);  // doubles per buffer: [TS][DIM] | [TS][QP] | [TS]
  const int num_mc = prm.num_mc;
  const int ntiles = (num_mc + TS - 1) / TS;
  auto load_tile = [&](int tile, int buf) {
    double* sx = sbuf + buf * kBuf;
    double* sc = sx + TS * DIM;
    double* sh = sc + TS * QP;
    const int i0 = tile * TS;
    for (int e = threadIdx.x; e < TS * DIM / 2; e += blockDim.x) {
      const bool ok = i0 + (2 * e) / DIM < num_mc;
      cp_async16(sx + 2 * e, ok ? xs + static_cast<size_t>(i0) * DIM + 2 * e : xs, ok);
    }
    for (int e = threadIdx.x; e < TS * QP / 2; e += blockDim.x) {
      const bool ok = i0 + (2 * e) / QP < num_mc;
      cp_async16(sc + 2 * e, ok ? cs + static_cast<size_t>(i0) * QP + 2 * e : cs, ok);
    }
    for (int e = threadIdx.x; e < TS; e += blockDim.x) {
      const bool ok = i0 + e < num_mc;
      cp_async8(sh + e, ok ? hs + i0 + e : hs, ok);
    }
  };
  load_tile(0, 0);
  cp_async_commit();
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) load_tile(tile + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const double* sx = sbuf + buf * kBuf;
    const double* sc = sx + TS * DIM;
    const double* sh = sc + TS * QP;
    const int cnt = min(TS, num_mc - tile * TS);
#pragma unroll 2
    for (int i = 0; i < cnt; ++i) {
      double xi[DIM], ci[QP];
#pragma unroll
      for (int d = 0; d < DIM; d += 2) {
        const double2 v = *reinterpret_cast<const double2*>(sx + i * DIM + d);
        xi[d] = v.x;
        xi[d + 1] = v.y;
      }
#pragma unroll
      for (int a = 0; a < QP; a += 2) {
        const double2 v = *reinterpret_cast<const double2*>(sc + i * QP + a);
        ci[a] = v.x;
        ci[a + 1] = v.y;
      }
      double dot = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) dot = fma(xi[d], xr[d], dot);
      double kv, kb;
      kernel_pair<KERNEL, TAB>(dot, pk0, sh[i], prm.alpha, kv, kb);
      if (GEN && mrow > 0) kv = kb * (sx[i * DIM + tcoord] - xrt) * tscale;
#pragma unroll
      for (int a = 0; a < QP; ++a) acc[a] = fma(ci[a], kv, acc[a]);
      if (!GEN && is_u) {
        double cu = 0.0;
#pragma unroll
        for (int a = 0; a < QP; ++a)
          if (a == uu) cu = ci[a];
        const double w = cu * kb;
        gkb += w;
#pragma unroll
        for (int d = 0; d < DIM; ++d) gu[d] = fma(w, xi[d], gu[d]);
      }
    }
    __syncthreads();  // the buffer is refilled two tiles later
  }
  if (!active) return;
  double* R = prm.R + static_cast<size_t>(cand) * QP * nrows;
#pragma unroll
  for (int a = 0; a < QP; ++a) R[static_cast<size_t>(a) * nrows + grow] = acc[a];
  if (!GEN && is_u) {
    prm.GkB[static_cast<size_t>(cand) * U + uu] = gkb;
#pragma unroll
    for (int d = 0; d < DIM; ++d) prm.Gu[(static_cast<size_t>(cand) * U + uu) * DIM + d] = gu[d];
  }
}

template <int KERNEL, int DIM, int QP>
__global__ void __launch_bounds__(128) kg_acc_kernel(const __grid_constant__ KgAccParams prm) {
  extern __shared__ __align__(16) double acc_smem[];
  if (prm.fast_exp) {
    exp_table_stage();
    kg_acc_body<KERNEL, DIM, QP, true>(prm, acc_smem);
    return;
  }
  kg_acc_body<KERNEL, DIM, QP, false>(prm, acc_smem);
}

template <int KERNEL, int DIM, int QP>
__global__ void __launch_bounds__(128) kg_acc_gen_kernel(const __grid_constant__ KgAccParams prm) {
  extern __shared__ __align__(16) double acc_smem[];
  if (prm.fast_exp) {
    exp_table_stage();
    kg_acc_body<KERNEL, DIM, QP, true, true>(prm, acc_smem);
    return;
  }
  kg_acc_body<KERNEL, DIM, QP, false, true>(prm, acc_smem);
}

// launchers instantiated per (DIM) translation unit
using KgMcLaunch = void (*)(const KgMcParams&, dim3 grid, size_t smem, cudaStream_t s);
using KgAccLaunch = void (*)(const KgAccParams&, dim3 grid, cudaStream_t s);
struct KgDispatchEntry {
  int kernel, dim, qp;
  KgMcLaunch mc;      // g == 0 fast path (TMA-staged operands)
  KgMcLaunch mc_gen;  // general path (derivative observations); may be null
  KgAccLaunch acc;
  KgAccLaunch acc_gen;  // general path; null where mc_gen is
  size_t (*smem_bytes)(int N, int U);
};
void register_kg_entries(const KgDispatchEntry* entries, int count);
const KgDispatchEntry* find_kg_entry(int kernel, int dim, int Q, bool need_gen);

template <int KERNEL, int DIM, int QP>
void launch_kg_mc(const KgMcParams& p, dim3 grid, size_t smem, cudaStream_t s) {
  // function attributes are per device: set on every launch (cheap), never cached process-wide
  cudaFuncSetAttribute(kg_mc_kernel<KERNEL, DIM, QP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  kg_mc_kernel<KERNEL, DIM, QP><<<grid, kMcThreads, smem, s>>>(p);
}
template <int KERNEL, int DIM, int QP>
void launch_kg_mc_gen(const KgMcParams& p, dim3, size_t, cudaStream_t s) {
  int dev = 0, sms = 1, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t ring_bytes = static_cast<size_t>(kRingRows + kRingDepth) * kMcThreads * sizeof(double) +
                            (p.stage_ops ? static_cast<size_t>(p.N) * (DIM + 2) * sizeof(double) : 0);
  cudaFuncSetAttribute(kg_mc_gen_kernel<KERNEL, DIM, QP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       static_cast<int>(ring_bytes));
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kg_mc_gen_kernel<KERNEL, DIM, QP>, kMcThreads, ring_bytes);
  const int grid = std::max(1, std::min(std::min(p.aw_slots, p.work_total), sms * std::max(1, per_sm)));
  kg_mc_gen_kernel<KERNEL, DIM, QP><<<grid, kMcThreads, ring_bytes, s>>>(p);
}
template <int KERNEL, int DIM, int QP>
void launch_kg_acc(const KgAccParams& p, dim3 grid, cudaStream_t s) {
  const size_t smem = 2 * static_cast<size_t>(kAccTile) * (DIM + QP + 1) * sizeof(double);
  cudaFuncSetAttribute(kg_acc_kernel<KERNEL, DIM, QP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  kg_acc_kernel<KERNEL, DIM, QP><<<grid, 128, smem, s>>>(p);
}
template <int KERNEL, int DIM, int QP>
void launch_kg_acc_gen(const KgAccParams& p, dim3 grid, cudaStream_t s) {
  const size_t smem = 2 * static_cast<size_t>(kAccTile) * (DIM + QP + 1) * sizeof(double);
  cudaFuncSetAttribute(kg_acc_gen_kernel<KERNEL, DIM, QP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  kg_acc_gen_kernel<KERNEL, DIM, QP><<<grid, 128, smem, s>>>(p);
}
template <int DIM, int QP>
size_t kg_smem_bytes(int N, int U) {
  return (static_cast<size_t>(N) * (DIM + QP + 2) + static_cast<size_t>(U) * (DIM + 2)) * sizeof(double);
}

#define CMOE_KG_ENTRY(K, D, Q) \
  { K, D, Q, &launch_kg_mc<K, D, Q>, nullptr, &launch_kg_acc<K, D, Q>, nullptr, &kg_smem_bytes<D, Q> }
#define CMOE_KG_ENTRY_GEN(K, D, Q)                                                                                \
  {                                                                                                               \
    K, D, Q, &launch_kg_mc<K, D, Q>, &launch_kg_mc_gen<K, D, Q>, &launch_kg_acc<K, D, Q>, &launch_kg_acc_gen<K, D, Q>, \
        &kg_smem_bytes<D, Q>                                                                                       \
  }
// union-row counts: 2, 4, 8, 16, 24, 32 (24 serves config 4: q = 4 with 4 derivative observations -> 20 rows; the
// 32-wide instantiation spilled ~1.5 KB per thread there)
#define CMOE_KG_ENTRIES_FOR_DIM(D)                                                                              \
  CMOE_KG_ENTRY(0, D, 2), CMOE_KG_ENTRY(0, D, 4), CMOE_KG_ENTRY_GEN(0, D, 8), CMOE_KG_ENTRY_GEN(0, D, 16),       \
      CMOE_KG_ENTRY_GEN(0, D, 24), CMOE_KG_ENTRY_GEN(0, D, 32), CMOE_KG_ENTRY(1, D, 2), CMOE_KG_ENTRY(1, D, 4),  \
      CMOE_KG_ENTRY_GEN(1, D, 8), CMOE_KG_ENTRY_GEN(1, D, 16), CMOE_KG_ENTRY_GEN(1, D, 24),                    \
      CMOE_KG_ENTRY_GEN(1, D, 32)

}  // namespace cmoe
