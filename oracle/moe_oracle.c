/* TEST INFRASTRUCTURE — the oracle is a checker, never the product (see moe_oracle.h).
 *
 * Plain-C restatement of the reference algorithm.  "ref:" comments cite
 * /root/reference/moe/optimal_learning/cpp/<file>:<lines>.  The arithmetic follows the reference's
 * formulation (augmented-GP re-solve per KG sample, per-sample gradient contraction), NOT the
 * rank-Q / accumulated formulation the CUDA path uses — that is the point of a checker.
 */
#include "moe_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#define SQRT5 2.236067977499789696409173668731276235440618359611525724270897

struct oracle_gp {
  int kernel, dim, N, g, n;
  double alpha;
  double* lengths_sq;
  double* X;      /* [N][dim] */
  double* y;      /* [N*(1+g)] interleaved */
  double* noise;  /* [1+g], indexed by observation TYPE (ref: gpp_math.cpp:447-449) */
  int* derivs;    /* [g] */
  double* K_chol; /* [n*n] col-major, lower */
  double* K_inv_y;
  double mean;
};

static void* xmalloc(size_t bytes) {
  void* p = malloc(bytes ? bytes : 1);
  if (!p) abort();
  return p;
}
static double* dalloc(size_t count) { return (double*)xmalloc(count * sizeof(double)); }
static double* dzero(size_t count) {
  double* p = dalloc(count);
  memset(p, 0, (count ? count : 1) * sizeof(double));
  return p;
}

/* ------------------------------------------------------------------------------------------------
 * Covariance kernels.  ref: gpp_covariance.cpp:121-164 (SE value block), :171-234 (SE grad block),
 * :339-387 (Matern value block), :389-459 (Matern grad block).
 * Block layout: cov[m + n*(1+g1)], m over point one's rows (0 = value, 1.. = d/dx_{d1[m-1]}),
 * n over point two's rows.  With u_a = (p2_a - p1_a)/l_a^2 the block is
 *   c00 = A ; c_m0 = B u_a ; c_0n = -B u_b ; c_mn = -C u_a u_b + [a==b] B / l_a^2
 * where (A,B,C) = (k,k,k) for SE and (cov0, first_derivative_part, alpha_exp_part) for Matern-5/2.
 * ---------------------------------------------------------------------------------------------- */
static double weighted_sqdist(const double* a, const double* b, const double* lsq, int dim) {
  double norm = 0.0;
  for (int i = 0; i < dim; ++i) {
    const double diff = a[i] - b[i];
    norm += diff * diff / lsq[i]; /* ref: gpp_covariance.cpp:56-58 */
  }
  return norm;
}

static void kernel_parts(int kernel, double alpha, double r2, double* A, double* B, double* C) {
  if (kernel == 0) {
    const double k = alpha * exp(-0.5 * r2);
    *A = k;
    *B = k;
    *C = k;
  } else {
    const double arg = SQRT5 * sqrt(r2);
    const double e = exp(-arg);
    *A = alpha * e * (1.0 + arg + 5.0 / 3.0 * r2);
    *B = 5.0 / 3.0 * alpha * e * (arg + 1.0);
    *C = 25.0 / 3.0 * alpha * e;
  }
}

static void cov_block(int kernel, int dim, double alpha, const double* lsq, const double* p1, const int* d1, int g1,
                      const double* p2, const int* d2, int g2, double* cov) {
  const double r2 = weighted_sqdist(p1, p2, lsq, dim);
  double A, B, C;
  kernel_parts(kernel, alpha, r2, &A, &B, &C);
  const int ld = 1 + g1;
  cov[0] = A;
  for (int m = 0; m < g1; ++m) {
    const int a = d1[m];
    cov[m + 1] = B * ((p2[a] - p1[a]) / lsq[a]);
  }
  for (int n = 0; n < g2; ++n) {
    const int b = d2[n];
    cov[(n + 1) * ld] = B * ((p1[b] - p2[b]) / lsq[b]);
  }
  for (int m = 0; m < g1; ++m) {
    const int a = d1[m];
    const double ua = (p2[a] - p1[a]) / lsq[a];
    for (int n = 0; n < g2; ++n) {
      const int b = d2[n];
      const double vb = (p1[b] - p2[b]) / lsq[b];
      double v = ua * vb * C;
      if (a == b) v += B / lsq[b];
      cov[(m + 1) + (n + 1) * ld] = v;
    }
  }
}

/* gradient of the block wrt point one: grad[i + m*dim + n*dim*(1+g1)] */
static void grad_cov_block(int kernel, int dim, double alpha, const double* lsq, const double* p1, const int* d1,
                           int g1, const double* p2, const int* d2, int g2, double* grad) {
  const double r2 = weighted_sqdist(p1, p2, lsq, dim);
  double A, B, C;
  kernel_parts(kernel, alpha, r2, &A, &B, &C);
  (void)A;
  const int ld = 1 + g1;
  for (int i = 0; i < dim; ++i) {
    const double ui = (p2[i] - p1[i]) / lsq[i];
    grad[i] = ui * B;
    for (int m = 0; m < g1; ++m) {
      const int a = d1[m];
      const double ua = (p2[a] - p1[a]) / lsq[a];
      double v = C * ui * ua;
      if (i == a) v -= B / lsq[a];
      grad[i + (m + 1) * dim] = v;
    }
    for (int n = 0; n < g2; ++n) {
      const int b = d2[n];
      const double vb = (p1[b] - p2[b]) / lsq[b];
      double v = C * ui * vb;
      if (i == b) v += B / lsq[b];
      grad[i + (n + 1) * dim * ld] = v;
    }
    for (int m = 0; m < g1; ++m) {
      const int a = d1[m];
      const double ua = (p2[a] - p1[a]) / lsq[a];
      for (int n = 0; n < g2; ++n) {
        const int b = d2[n];
        const double vb = (p1[b] - p2[b]) / lsq[b];
        double v;
        if (kernel == 0) {
          /* ref: gpp_covariance.cpp:219-230 */
          v = ua * vb;
          if (a == b) v += 1.0 / lsq[a];
          v *= ui;
          if (a == i) v -= vb / lsq[a];
          if (b == i) v += ua / lsq[b];
          v *= C;
        } else if (r2 > 0.0) {
          /* ref: gpp_covariance.cpp:441-452 */
          v = C * ua * vb;
          v *= SQRT5 * ui / sqrt(r2);
          if (a == i) v -= C * vb / lsq[a];
          if (b == i) v += C * ua / lsq[b];
          if (a == b) v += C * ui / lsq[a];
        } else {
          v = 0.0; /* ref: gpp_covariance.cpp:453-455 */
        }
        grad[i + (m + 1) * dim + (n + 1) * dim * ld] = v;
      }
    }
  }
}

static double* lengths_squared(const double* lengths, int dim) {
  double* lsq = dalloc(dim);
  for (int i = 0; i < dim; ++i) lsq[i] = lengths[i] * lengths[i];
  return lsq;
}

void oracle_covariance(int kernel, int dim, double alpha, const double* lengths, const double* p1, const int* d1,
                       int g1, const double* p2, const int* d2, int g2, double* cov) {
  double* lsq = lengths_squared(lengths, dim);
  cov_block(kernel, dim, alpha, lsq, p1, d1, g1, p2, d2, g2, cov);
  free(lsq);
}

void oracle_grad_covariance(int kernel, int dim, double alpha, const double* lengths, const double* p1, const int* d1,
                            int g1, const double* p2, const int* d2, int g2, double* grad_cov) {
  double* lsq = lengths_squared(lengths, dim);
  grad_cov_block(kernel, dim, alpha, lsq, p1, d1, g1, p2, d2, g2, grad_cov);
  free(lsq);
}

/* ------------------------------------------------------------------------------------------------
 * Dense linear algebra.  ref: gpp_linear_algebra.cpp:109-148 (Cholesky, pivot test > 1e-16, returns k+1),
 * :160-193 (triangular solve), :203-208 (multi-RHS), gpp_linear_algebra.hpp:220,247 (potrs).
 * ---------------------------------------------------------------------------------------------- */
int oracle_cholesky(int n, double* a) {
  for (int k = 0; k < n; ++k) {
    double* colk = a + (size_t)k * n;
    if (!(colk[k] > 1.0e-16)) return k + 1;
    const double lkk = sqrt(colk[k]);
    colk[k] = lkk;
    for (int j = k + 1; j < n; ++j) colk[j] /= lkk;
    for (int j = k + 1; j < n; ++j) {
      double* colj = a + (size_t)j * n;
      const double ljk = colk[j];
      for (int i = j; i < n; ++i) colj[i] = colj[i] - colk[i] * ljk;
    }
  }
  return 0;
}

void oracle_trsv(const double* a, int trans, int n, int lda, double* x) {
  if (!trans) {
    for (int j = 0; j < n; ++j) {
      const double* col = a + (size_t)j * lda;
      if (x[j] != 0.0) {
        x[j] /= col[j];
        const double t = x[j];
        for (int i = j + 1; i < n; ++i) x[i] = x[i] - t * col[i];
      }
    }
  } else {
    for (int j = n - 1; j >= 0; --j) {
      const double* col = a + (size_t)j * lda;
      double t = x[j];
      for (int i = n - 1; i >= j + 1; --i) t -= col[i] * x[i];
      x[j] = t / col[j];
    }
  }
}

static void trsm(const double* a, int trans, int n, int nrhs, int lda, double* x) {
  for (int k = 0; k < nrhs; ++k) oracle_trsv(a, trans, n, lda, x + (size_t)k * n);
}

void oracle_potrs(const double* a, int n, int nrhs, double* x) {
  trsm(a, 0, n, nrhs, n, x);
  trsm(a, 1, n, nrhs, n, x);
}

/* y(+)= alpha * op(A) x ; ref: gpp_linear_algebra.cpp:340-372 */
static void gemv(const double* a, int trans, const double* x, double alpha, double beta, int m, int ncol, int lda,
                 double* y) {
  const int leny = trans ? ncol : m;
  if (beta == 0.0) {
    for (int i = 0; i < leny; ++i) y[i] = 0.0;
  } else if (beta != 1.0) {
    for (int i = 0; i < leny; ++i) y[i] *= beta;
  }
  if (!trans) {
    for (int i = 0; i < ncol; ++i) {
      const double t = alpha * x[i];
      const double* col = a + (size_t)i * lda;
      for (int j = 0; j < m; ++j) y[j] += col[j] * t;
    }
  } else {
    for (int i = 0; i < ncol; ++i) {
      const double* col = a + (size_t)i * lda;
      double t = 0.0;
      for (int j = 0; j < m; ++j) t += col[j] * x[j];
      y[i] += alpha * t;
    }
  }
}

/* C = alpha*op(A)*B + beta*C ; ref: gpp_linear_algebra.cpp:384-398 (a loop of GEMVs) */
static void gemm(const double* a, int transa, const double* b, double alpha, double beta, int m, int k, int ncol,
                 double* c) {
  for (int j = 0; j < ncol; ++j) {
    if (!transa) {
      gemv(a, 0, b + (size_t)j * k, alpha, beta, m, k, m, c + (size_t)j * m);
    } else {
      gemv(a, 1, b + (size_t)j * k, alpha, beta, k, m, k, c + (size_t)j * m);
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * GP construction.  ref: gpp_math.cpp:426-455 (K + noise by observation type, lower triangle),
 * :481-511 (Cholesky, mean_ = average of function values, K^-1 (y - mean)), :553-573 (ctor).
 * ---------------------------------------------------------------------------------------------- */
static void build_K_with_noise(int kernel, int dim, double alpha, const double* lsq, const double* noise,
                               const double* X, int N, const int* derivs, int g, double* K) {
  const int b = 1 + g, n = N * b;
  double* blk = dalloc((size_t)b * b);
  for (int i = 0; i < N; ++i) {
    for (int j = i; j < N; ++j) {
      cov_block(kernel, dim, alpha, lsq, X + (size_t)j * dim, derivs, g, X + (size_t)i * dim, derivs, g, blk);
      for (int m = 0; m < b; ++m) {
        for (int c = 0; c < b; ++c) {
          const int row = j * b + m, col = i * b + c;
          if (row >= col) K[row + (size_t)col * n] = blk[m + c * b];
          if (row == col) K[row + (size_t)col * n] += noise[m];
        }
      }
    }
  }
  free(blk);
}

/* K(X, pts) with derivative rows on both sides; ref: gpp_math.cpp:309-335 */
static void build_mix_cov(int kernel, int dim, double alpha, const double* lsq, const double* X, int N,
                          const int* dX, int gX, const double* pts, int num, const int* dP, int gP, double* out) {
  const int bx = 1 + gX, bp = 1 + gP, rows = N * bx;
  double* blk = dalloc((size_t)bx * bp);
  for (int j = 0; j < num; ++j) {
    for (int i = 0; i < N; ++i) {
      cov_block(kernel, dim, alpha, lsq, X + (size_t)i * dim, dX, gX, pts + (size_t)j * dim, dP, gP, blk);
      for (int m = 0; m < bx; ++m)
        for (int c = 0; c < bp; ++c) out[(i * bx + m) + (size_t)(j * bp + c) * rows] = blk[m + c * bx];
    }
  }
  free(blk);
}

static int gp_refit(oracle_gp* gp, int mean_change) {
  const int n = gp->N * (1 + gp->g);
  gp->n = n;
  build_K_with_noise(gp->kernel, gp->dim, gp->alpha, gp->lengths_sq, gp->noise, gp->X, gp->N, gp->derivs, gp->g,
                     gp->K_chol);
  const int lm = oracle_cholesky(n, gp->K_chol);
  if (lm != 0) return lm;
  if (mean_change) {
    double s = 0.0;
    for (int i = 0; i < gp->N; ++i) s += gp->y[(size_t)i * (1 + gp->g)];
    gp->mean = s / gp->N;
  }
  memcpy(gp->K_inv_y, gp->y, (size_t)n * sizeof(double));
  for (int i = 0; i < gp->N; ++i) gp->K_inv_y[(size_t)i * (1 + gp->g)] -= gp->mean;
  oracle_potrs(gp->K_chol, n, 1, gp->K_inv_y);
  return 0;
}

oracle_gp* oracle_gp_create(int kernel, double alpha, const double* lengths, const double* X, const double* y,
                            const double* noise, const int* derivs, int g, int dim, int N, int* leading_minor) {
  oracle_gp* gp = (oracle_gp*)xmalloc(sizeof(oracle_gp));
  const int n = N * (1 + g);
  gp->kernel = kernel;
  gp->dim = dim;
  gp->N = N;
  gp->g = g;
  gp->n = n;
  gp->alpha = alpha;
  gp->lengths_sq = lengths_squared(lengths, dim);
  gp->X = dalloc((size_t)N * dim);
  memcpy(gp->X, X, (size_t)N * dim * sizeof(double));
  gp->y = dalloc(n);
  memcpy(gp->y, y, (size_t)n * sizeof(double));
  gp->noise = dalloc(1 + g);
  memcpy(gp->noise, noise, (size_t)(1 + g) * sizeof(double));
  gp->derivs = (int*)xmalloc((size_t)(g ? g : 1) * sizeof(int));
  if (g) memcpy(gp->derivs, derivs, (size_t)g * sizeof(int));
  gp->K_chol = dzero((size_t)n * n);
  gp->K_inv_y = dalloc(n);
  gp->mean = 0.0;
  *leading_minor = gp_refit(gp, 1);
  if (*leading_minor != 0) {
    oracle_gp_destroy(gp);
    return NULL;
  }
  return gp;
}

void oracle_gp_destroy(oracle_gp* gp) {
  if (!gp) return;
  free(gp->lengths_sq);
  free(gp->X);
  free(gp->y);
  free(gp->noise);
  free(gp->derivs);
  free(gp->K_chol);
  free(gp->K_inv_y);
  free(gp);
}

void oracle_gp_get_state(const oracle_gp* gp, double* K_chol, double* K_inv_y, double* mean) {
  if (K_chol) memcpy(K_chol, gp->K_chol, (size_t)gp->n * gp->n * sizeof(double));
  if (K_inv_y) memcpy(K_inv_y, gp->K_inv_y, (size_t)gp->n * sizeof(double));
  if (mean) *mean = gp->mean;
}

/* ------------------------------------------------------------------------------------------------
 * Per-point-set cache.  ref: PointsToSampleState, gpp_math.hpp:889-987; FillPointsToSampleState,
 * gpp_math.cpp:600-653.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int num, gs, Q, nd; /* nd = number of leading points we differentiate against */
  const int* ds;
  const double* pts;
  double* K_star;      /* [n][Q] */
  double* KinvKs;      /* [n][Q] */
  double* grad_K_star; /* [nd*(1+gs)][n][dim], d fastest */
} pstate;

static void pstate_init(pstate* st, const oracle_gp* gp, const double* pts, int num, const int* ds, int gs, int nd) {
  const int n = gp->n, dim = gp->dim, bs = 1 + gs, bx = 1 + gp->g;
  st->num = num;
  st->gs = gs;
  st->Q = num * bs;
  st->nd = nd;
  st->ds = ds;
  st->pts = pts;
  st->K_star = dalloc((size_t)n * st->Q);
  st->KinvKs = dalloc((size_t)n * st->Q);
  build_mix_cov(gp->kernel, dim, gp->alpha, gp->lengths_sq, gp->X, gp->N, gp->derivs, gp->g, pts, num, ds, gs,
                st->K_star);
  memcpy(st->KinvKs, st->K_star, (size_t)n * st->Q * sizeof(double));
  oracle_potrs(gp->K_chol, n, st->Q, st->KinvKs);
  st->grad_K_star = NULL;
  if (nd > 0) {
    st->grad_K_star = dalloc((size_t)nd * bs * n * dim);
    double* blk = dalloc((size_t)dim * bs * bx);
    for (int i = 0; i < nd; ++i) {
      for (int j = 0; j < gp->N; ++j) {
        grad_cov_block(gp->kernel, dim, gp->alpha, gp->lengths_sq, pts + (size_t)i * dim, ds, gs,
                       gp->X + (size_t)j * dim, gp->derivs, gp->g, blk);
        for (int m = 0; m < bs; ++m)
          for (int c = 0; c < bx; ++c) {
            const int row = c + j * bx, col = m + i * bs;
            for (int d = 0; d < dim; ++d)
              st->grad_K_star[d + (size_t)row * dim + (size_t)col * dim * n] = blk[d + m * dim + c * dim * bs];
          }
      }
    }
    free(blk);
  }
}

static void pstate_free(pstate* st) {
  free(st->K_star);
  free(st->KinvKs);
  free(st->grad_K_star);
}

/* mu = mean_ (value rows) + K*^T K^-1 y ; ref: gpp_math.cpp:662-678 */
static void gp_mean(const oracle_gp* gp, const pstate* st, double* mu) {
  for (int i = 0; i < st->num; ++i)
    for (int j = 0; j < 1 + st->gs; ++j) mu[i * (1 + st->gs) + j] = (j == 0) ? gp->mean : 0.0;
  gemv(st->K_star, 1, gp->K_inv_y, 1.0, 1.0, gp->n, st->Q, gp->n, mu);
}

/* ref: gpp_math.cpp:688-710 */
void oracle_gp_mean_additional(const oracle_gp* gp, const double* pts, int num, double* mean) {
  double* kt = dalloc((size_t)gp->n * num);
  build_mix_cov(gp->kernel, gp->dim, gp->alpha, gp->lengths_sq, gp->X, gp->N, gp->derivs, gp->g, pts, num, NULL, 0,
                kt);
  for (int i = 0; i < num; ++i) mean[i] = gp->mean;
  gemv(kt, 1, gp->K_inv_y, 1.0, 1.0, gp->n, num, gp->n, mean);
  free(kt);
}

/* grad_mu[d + col*dim] = sum_row grad_K_star[d,row,col] K_inv_y[row] ; ref: gpp_math.cpp:721-726, 357-366 */
static void gp_grad_mean(const oracle_gp* gp, const pstate* st, double* grad_mu) {
  const int cols = st->nd * (1 + st->gs);
  for (int c = 0; c < cols; ++c)
    gemv(st->grad_K_star + (size_t)c * gp->dim * gp->n, 0, gp->K_inv_y, 1.0, 0.0, gp->dim, gp->n, gp->dim,
         grad_mu + (size_t)c * gp->dim);
}

/* Var = K(Xs,Xs) - (K^-1 K*)^T K* ; ref: gpp_math.cpp:924-970 (precomputed branch) */
static void gp_variance(const oracle_gp* gp, const pstate* st, double* var) {
  build_mix_cov(gp->kernel, gp->dim, gp->alpha, gp->lengths_sq, st->pts, st->num, st->ds, st->gs, st->pts, st->num,
                st->ds, st->gs, var);
  gemm(st->KinvKs, 1, st->K_star, -1.0, 1.0, st->Q, gp->n, st->Q, var);
}

/* d Var / d Xs_p ; output grad_var[d + row*dim + col*dim*Q].  ref: gpp_math.cpp:1267-1358.
 * Only block row/col p is non-zero; the (p,p) block gets both contributions. */
static void gp_grad_variance_point(const oracle_gp* gp, const pstate* st, int p, double* gv) {
  const int dim = gp->dim, n = gp->n, Q = st->Q, bs = 1 + st->gs;
  memset(gv, 0, (size_t)dim * Q * Q * sizeof(double));
  /* block column p: -(grad K*_{.,col})^T (K^-1 K*) for every row */
  for (int i = 0; i < bs; ++i) {
    const int col = p * bs + i;
    double* target = gv + (size_t)dim * Q * col;
    gemm(st->grad_K_star + (size_t)col * dim * n, 0, st->KinvKs, 1.0, 0.0, dim, n, Q, target);
    for (int e = 0; e < dim * Q; ++e) target[e] *= -1.0;
  }
  /* symmetrise the diagonal block: (p,m),(p,n) += transpose */
  for (int m = 0; m < bs; ++m)
    for (int c = m; c < bs; ++c)
      for (int d = 0; d < dim; ++d) {
        const size_t a = d + (size_t)(p * bs + m) * dim + (size_t)(p * bs + c) * dim * Q;
        const size_t b = d + (size_t)(p * bs + c) * dim + (size_t)(p * bs + m) * dim * Q;
        gv[a] += gv[b];
        gv[b] = gv[a];
      }
  /* leading term d K(Xs_p, Xs_j) / d Xs_p */
  double* blk = dalloc((size_t)dim * bs * bs);
  for (int j = 0; j < st->num; ++j) {
    grad_cov_block(gp->kernel, dim, gp->alpha, gp->lengths_sq, st->pts + (size_t)p * dim, st->ds, st->gs,
                   st->pts + (size_t)j * dim, st->ds, st->gs, blk);
    for (int m = 0; m < bs; ++m)
      for (int c = 0; c < bs; ++c) {
        const int row = j * bs + m, col = p * bs + c;
        for (int d = 0; d < dim; ++d) {
          double add = blk[d + c * dim + m * dim * bs];
          if (j == p) add += blk[d + m * dim + c * dim * bs];
          gv[d + (size_t)row * dim + (size_t)col * dim * Q] += add;
        }
      }
  }
  free(blk);
  /* mirror block column p into block row p */
  for (int i = 0; i < bs; ++i) {
    const int row = p * bs + i;
    for (int j = 0; j < st->num; ++j) {
      if (j == p) continue;
      for (int c = 0; c < bs; ++c) {
        const int col = j * bs + c;
        for (int d = 0; d < dim; ++d)
          gv[d + (size_t)row * dim + (size_t)col * dim * Q] = gv[d + (size_t)col * dim + (size_t)row * dim * Q];
      }
    }
  }
}

/* Smith (1995) forward differentiation of the outer-product Cholesky, in place on grad-variance.
 * On exit GL(d, k, j) (stored at [j*Q*dim + k*dim + d], j >= k) = d L_{jk}.  ref: gpp_math.cpp:1389-1458. */
static void gp_grad_chol_point(const oracle_gp* gp, const pstate* st, int p, const double* chol, double* gc) {
  const int dim = gp->dim, Q = st->Q;
  gp_grad_variance_point(gp, st, p, gc);
  for (int i = 0; i < Q; ++i)
    for (int j = (i + 1) * dim; j < dim * Q; ++j) gc[(size_t)i * Q * dim + j] = 0.0;
#define L_(i, j) chol[(size_t)(j)*Q + (i)]
#define G_(m, i, j) gc[(size_t)(j)*Q * dim + (size_t)(i)*dim + (m)]
  const double eps = 2.220446049250313e-16; /* kMinimumStdDev, gpp_math.hpp:291 */
  for (int k = 0; k < Q; ++k) {
    const double lkk = L_(k, k);
    if (lkk > eps) {
      for (int m = 0; m < dim; ++m) G_(m, k, k) = 0.5 * G_(m, k, k) / lkk;
      for (int j = k + 1; j < Q; ++j)
        for (int m = 0; m < dim; ++m) G_(m, k, j) = (G_(m, k, j) - L_(j, k) * G_(m, k, k)) / lkk;
      for (int j = k + 1; j < Q; ++j)
        for (int i = j; i < Q; ++i)
          for (int m = 0; m < dim; ++m)
            G_(m, j, i) = G_(m, j, i) - G_(m, k, i) * L_(j, k) - L_(i, k) * G_(m, k, j);
    }
  }
#undef L_
#undef G_
}

int oracle_gp_posterior(const oracle_gp* gp, const double* pts, int num, const int* derivs_s, int g_s, double* mean,
                        double* grad_mean, double* var, double* chol_var, double* grad_var, double* grad_chol) {
  const int need_grad = (grad_mean || grad_var || grad_chol);
  pstate st;
  pstate_init(&st, gp, pts, num, derivs_s, g_s, need_grad ? num : 0);
  const int Q = st.Q;
  int rc = 0;
  if (mean) gp_mean(gp, &st, mean);
  if (grad_mean) gp_grad_mean(gp, &st, grad_mean);
  double* v = dalloc((size_t)Q * Q);
  gp_variance(gp, &st, v);
  if (var) memcpy(var, v, (size_t)Q * Q * sizeof(double));
  const size_t blk = (size_t)gp->dim * Q * Q;
  if (grad_var)
    for (int p = 0; p < num; ++p) gp_grad_variance_point(gp, &st, p, grad_var + blk * p);
  if (chol_var || grad_chol) {
    rc = oracle_cholesky(Q, v);
    if (rc == 0) {
      if (chol_var) memcpy(chol_var, v, (size_t)Q * Q * sizeof(double));
      if (grad_chol)
        for (int p = 0; p < num; ++p) gp_grad_chol_point(gp, &st, p, v, grad_chol + blk * p);
    }
  }
  free(v);
  pstate_free(&st);
  return rc;
}

/* ------------------------------------------------------------------------------------------------
 * q-EI Monte Carlo.  ref: gpp_math.cpp:1991-2033 (value), :2050-2126 (gradient).
 * Candidate points carry no derivative rows (gpp_math.cpp:2149-2150); jitter 1e-6 on the diagonal.
 * `next` draws one normal (table replay or Philox).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const double* table;
  int len, pos;
} normal_src;

static double draw_table(normal_src* s) {
  if (s->pos >= s->len) abort(); /* the reference throws InvalidValueException here (gpp_random.cpp:146-154) */
  return s->table[s->pos++];
}

double oracle_ei(const oracle_gp* gp, const double* Xq, const double* Xp, int q, int p, int num_mc,
                 double best_so_far, const double* table, int table_len, double* grad) {
  const int dim = gp->dim, U = q + p;
  double* un = dalloc((size_t)U * dim);
  memcpy(un, Xq, (size_t)q * dim * sizeof(double));
  if (p) memcpy(un + (size_t)q * dim, Xp, (size_t)p * dim * sizeof(double));
  pstate st;
  pstate_init(&st, gp, un, U, NULL, 0, grad ? q : 0);
  double* mu = dalloc(U);
  double* L = dalloc((size_t)U * U);
  gp_mean(gp, &st, mu);
  gp_variance(gp, &st, L);
  for (int i = 0; i < U; ++i) L[i + (size_t)i * U] += 1.0e-6;
  if (oracle_cholesky(U, L) != 0) {
    free(un); free(mu); free(L); pstate_free(&st);
    return NAN;
  }
  double* grad_mu = NULL;
  double* gchol = NULL;
  double* agg = NULL;
  if (grad) {
    grad_mu = dalloc((size_t)dim * q);
    gp_grad_mean(gp, &st, grad_mu);
    gchol = dalloc((size_t)dim * U * U * q);
    for (int k = 0; k < q; ++k) gp_grad_chol_point(gp, &st, k, L, gchol + (size_t)dim * U * U * k);
    agg = dzero((size_t)dim * q);
  }
  normal_src src = {table, table_len, 0};
  double* z = dalloc(U);
  double* lz = dalloc(U);
  double total = 0.0;
  for (int it = 0; it < num_mc; ++it) {
    for (int j = 0; j < U; ++j) z[j] = draw_table(&src);
    /* y = L z (lower triangle only; ref uses TriangularMatrixVectorMultiply, gpp_linear_algebra.cpp:257-287) */
    for (int i = 0; i < U; ++i) {
      double s = L[i + (size_t)i * U] * z[i]; /* same summation order as the in-place sweep of the reference */
      for (int j = i - 1; j >= 0; --j) s += L[i + (size_t)j * U] * z[j];
      lz[i] = s;
    }
    double imp = 0.0;
    int winner = U + 1;
    for (int j = 0; j < U; ++j) {
      const double e = best_so_far - (mu[j] + lz[j]);
      if (e > imp) {
        imp = e;
        winner = j;
      }
    }
    if (imp > 0.0) {
      total += imp;
      if (grad) {
        if (winner < q)
          for (int d = 0; d < dim; ++d) agg[winner * dim + d] -= grad_mu[winner * dim + d];
        /* agg[k] -= dL[:, :, winner, k] z ; ref: gpp_math.cpp:2114-2119 */
        for (int k = 0; k < q; ++k) {
          const double* blk = gchol + (size_t)dim * U * U * k + (size_t)winner * dim * U;
          for (int i = 0; i < U; ++i)
            for (int d = 0; d < dim; ++d) agg[k * dim + d] -= blk[d + (size_t)i * dim] * z[i];
        }
      }
    }
  }
  if (grad)
    for (int e = 0; e < q * dim; ++e) grad[e] = agg[e] / (double)num_mc;
  free(un); free(mu); free(L); free(grad_mu); free(gchol); free(agg); free(z); free(lz);
  pstate_free(&st);
  return total / (double)num_mc;
}

/* ------------------------------------------------------------------------------------------------
 * Domain step limiter.  ref: gpp_domain.cpp:64-104 (kInvalidStepScaleFactor = 0.5, gpp_domain.hpp).
 * ---------------------------------------------------------------------------------------------- */
void oracle_limit_update(const double* bounds, int dim, double mrc, const double* x, double* upd) {
  for (int j = 0; j < dim; ++j) {
    const double lo = bounds[2 * j], hi = bounds[2 * j + 1];
    double step = upd[j];
    double dist = fmin(x[j] - lo, hi - x[j]);
    if (fabs(step) > mrc * dist) step = copysign(mrc * dist, step);
    const double next = x[j] + step;
    if (next < lo || next > hi) {
      if (next < lo) {
        dist = lo - x[j];
        step = (x[j] + step * 0.5 < lo) ? dist * 0.5 : step * 0.5;
      } else {
        dist = hi - x[j];
        step = (x[j] + step * 0.5 > hi) ? dist * 0.5 : step * 0.5;
      }
    }
    upd[j] = step;
  }
}

/* ------------------------------------------------------------------------------------------------
 * q-KG / d-KG Monte Carlo.
 * ref: gpp_knowledge_gradient_optimization.cpp:69-115 (value), :130-227 (gradient), :246-317 (state / PreCompute),
 *      :420-472 (ComputeOptimalPosteriorMean), gpp_optimization.hpp:708-828 (line-search GD), :1242-1283 (restarts).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const oracle_gp* base;
  oracle_gp aug; /* augmented GP: X u union, K_chol over n+Q, mean frozen */
  int nf;        /* num_fidelity */
} kg_ctx;

/* posterior mean / gradient of the augmented GP at one full-dim point (value row only) */
static double aug_mean(const oracle_gp* a, const double* x, double* krow) {
  build_mix_cov(a->kernel, a->dim, a->alpha, a->lengths_sq, a->X, a->N, a->derivs, a->g, x, 1, NULL, 0, krow);
  double m = a->mean;
  double t = 0.0;
  for (int j = 0; j < a->n; ++j) t += krow[j] * a->K_inv_y[j];
  return m + 1.0 * t;
}

static void aug_grad_mean(const oracle_gp* a, const double* x, double* gm) {
  /* ref: gpp_math.cpp:728-757 */
  const int dim = a->dim, bx = 1 + a->g;
  double* blk = dalloc((size_t)dim * bx);
  for (int d = 0; d < dim; ++d) gm[d] = 0.0;
  /* the reference accumulates column by column of the [dim][n] tensor (GEMV 'N'), keep that order */
  for (int j = 0; j < a->N; ++j) {
    grad_cov_block(a->kernel, dim, a->alpha, a->lengths_sq, x, NULL, 0, a->X + (size_t)j * dim, a->derivs, a->g, blk);
    for (int c = 0; c < bx; ++c) {
      const double t = 1.0 * a->K_inv_y[c + j * bx];
      for (int d = 0; d < dim; ++d) gm[d] += blk[d + c * dim] * t;
    }
  }
  free(blk);
}

typedef struct {
  double* x; /* full-dim current point; fidelity coordinates pinned at 1.0 (ref: .cpp:361-366) */
  double* krow;
  double* g;
} pm_state;

/* objective = -mu ; ref: .cpp:334-359 */
static double pm_obj(const kg_ctx* c, pm_state* s) { return -aug_mean(&c->aug, s->x, s->krow); }
static void pm_grad(const kg_ctx* c, pm_state* s, double* out) {
  aug_grad_mean(&c->aug, s->x, s->g);
  for (int i = 0; i < c->aug.dim - c->nf; ++i) out[i] = -s->g[i];
}

/* Line-search statistics for sizing the device kernel (test infrastructure only):
 * [0..30] histogram of backtracking counts per step, [32..47] histogram of steps taken per descent,
 * [48] breaks on no-improvement, [49] breaks on 30 halvings, [50] breaks on the step tolerance, [51] ran all steps. */
static long g_ls_stats[64];
void oracle_debug_line_search_stats(long* out, int reset) {
  for (int i = 0; i < 64; ++i) {
    out[i] = g_ls_stats[i];
    if (reset) g_ls_stats[i] = 0;
  }
}
static void ls_count(int slot) {
#pragma omp atomic
  g_ls_stats[slot] += 1;
}

static void line_search_gd(const kg_ctx* c, const double* gd, const double* bounds, pm_state* s) {
  const int ps = c->aug.dim - c->nf;
  const int max_steps = (int)gd[1];
  const double gamma = gd[4], pre_mult = gd[5], mrc = gd[6], tol = gd[7];
  double* grad = dalloc(ps);
  double* step = dalloc(ps);
  double* trial = dalloc(ps);
  double* next = dalloc(ps);
  memcpy(next, s->x, (size_t)ps * sizeof(double));
  const double step_tol = tol / (double)max_steps;
  for (int i = 0; i < max_steps; ++i) {
    const double f0 = pm_obj(c, s);
    double obj = 0.0;
    double alpha = pre_mult * pow((double)(i + 1), -gamma);
    pm_grad(c, s, grad);
    double nrm = 0.0;
    for (int j = 0; j < ps; ++j) nrm += grad[j] * grad[j];
    int search = 0;
    while (search < 30) {
      for (int j = 0; j < ps; ++j) trial[j] = next[j] + alpha * grad[j];
      memcpy(s->x, trial, (size_t)ps * sizeof(double));
      obj = pm_obj(c, s);
      if (obj - f0 > 0.5 * alpha * nrm) break;
      alpha *= 0.5;
      search += 1;
    }
    for (int j = 0; j < ps; ++j) step[j] = alpha * grad[j];
    oracle_limit_update(bounds, ps, mrc, next, step);
    for (int j = 0; j < ps; ++j) trial[j] = next[j] + step[j];
    memcpy(s->x, trial, (size_t)ps * sizeof(double));
    obj = pm_obj(c, s);
    ls_count(search);
    if (obj <= f0 || search == 30) {
      memcpy(s->x, next, (size_t)ps * sizeof(double));
      ls_count(search == 30 ? 49 : 48);
      ls_count(32 + (i < 15 ? i : 15));
      break;
    }
    for (int j = 0; j < ps; ++j) next[j] += step[j];
    memcpy(s->x, next, (size_t)ps * sizeof(double));
    double ns = 0.0;
    for (int j = 0; j < ps; ++j) ns += step[j] * step[j];
    if (sqrt(ns) < step_tol) {
      ls_count(50);
      ls_count(32 + (i + 1 < 15 ? i + 1 : 15));
      break;
    }
    if (i == max_steps - 1) {
      ls_count(51);
      ls_count(32 + (i + 1 < 15 ? i + 1 : 15));
    }
  }
  free(grad); free(step); free(trial); free(next);
}

/* ref: .cpp:420-472.  Returns 0 and leaves outputs untouched when max_num_restarts <= 0. */
static int optimal_posterior_mean(const kg_ctx* c, const double* gd, const double* bounds, const double* starts,
                                  int num_starts, double* best_point, double* best_value) {
  const int restarts = (int)gd[2];
  if (restarts <= 0) return 0;
  const int dim = c->aug.dim, ps = dim - c->nf;
  pm_state s;
  s.x = dalloc(dim);
  s.krow = dalloc(c->aug.n);
  s.g = dalloc(dim);
  for (int d = ps; d < dim; ++d) s.x[d] = 1.0;
  int arg = 0;
  double best_mean = 0.0;
  for (int i = 0; i < num_starts; ++i) {
    memcpy(s.x, starts + (size_t)i * ps, (size_t)ps * sizeof(double));
    const double mval = -pm_obj(c, &s);
    if (i == 0 || best_mean > mval) {
      best_mean = mval;
      arg = i;
    }
  }
  memcpy(s.x, starts + (size_t)arg * ps, (size_t)ps * sizeof(double));
  double* cur = dalloc(ps);
  for (int r = 0; r < restarts; ++r) {
    memcpy(cur, s.x, (size_t)ps * sizeof(double));
    line_search_gd(c, gd, bounds, &s);
    double nd = 0.0;
    for (int j = 0; j < ps; ++j) nd += (cur[j] - s.x[j]) * (cur[j] - s.x[j]);
    if (sqrt(nd) <= gd[7]) break;
  }
  *best_value = pm_obj(c, &s);
  memcpy(best_point, s.x, (size_t)ps * sizeof(double));
  free(cur); free(s.x); free(s.krow); free(s.g);
  return 1;
}

static double kg_impl(const oracle_gp* gp, int nf, const double* gd, const double* inner_bounds,
                      const double* discrete_pts, int num_pts, const double* Xq, const double* Xp, int q, int p,
                      int num_mc, double best_so_far, normal_src* src, double* grad, double* best_points_out) {
  const int dim = gp->dim, g = gp->g, bs = 1 + g, U = q + p, Q = U * bs, n = gp->n, ps = dim - nf;
  double* un = dalloc((size_t)U * dim);
  memcpy(un, Xq, (size_t)q * dim * sizeof(double));
  if (p) memcpy(un + (size_t)q * dim, Xp, (size_t)p * dim * sizeof(double));
  /* discretised set = [union (non-fidelity coords) ; discrete_pts] ; ref: .cpp:259-261 */
  const int M = U + num_pts;
  double* dset = dalloc((size_t)M * ps);
  for (int i = 0; i < U; ++i) memcpy(dset + (size_t)i * ps, un + (size_t)i * dim, (size_t)ps * sizeof(double));
  memcpy(dset + (size_t)U * ps, discrete_pts, (size_t)num_pts * ps * sizeof(double));

  pstate st;
  pstate_init(&st, gp, un, U, gp->derivs, g, grad ? q : 0);
  double* mu = dalloc(Q);
  gp_mean(gp, &st, mu);
  double* L = dalloc((size_t)Q * Q);
  gp_variance(gp, &st, L);
  for (int i = 0; i < U; ++i)
    for (int j = 0; j < bs; ++j) {
      const int r = i * bs + j;
      L[r + (size_t)r * Q] += gp->noise[j]; /* ref: .cpp:304-309 */
    }
  double result = NAN;
  if (oracle_cholesky(Q, L) != 0) goto done_early;
  for (int j = 1; j < Q; ++j)
    for (int i = 0; i < j; ++i) L[i + (size_t)j * Q] = 0.0; /* ZeroUpperTriangle, ref: .cpp:316 */

  {
    int winner = -1;
    double best_post = best_so_far;
    for (int j = 0; j < U; ++j)
      if (mu[j * bs] < best_post) {
        best_post = mu[j * bs];
        winner = j;
      }
    /* augmented GP: K over X u union re-built and re-factored ONCE (ref: .cpp:83-85, gpp_math.cpp:1720-1737) */
    kg_ctx c;
    c.base = gp;
    c.nf = nf;
    c.aug = *gp;
    c.aug.N = gp->N + U;
    c.aug.n = c.aug.N * bs;
    c.aug.X = dalloc((size_t)c.aug.N * dim);
    memcpy(c.aug.X, gp->X, (size_t)gp->N * dim * sizeof(double));
    memcpy(c.aug.X + (size_t)gp->N * dim, un, (size_t)U * dim * sizeof(double));
    c.aug.y = dalloc(c.aug.n);
    memcpy(c.aug.y, gp->y, (size_t)n * sizeof(double));
    c.aug.K_chol = dzero((size_t)c.aug.n * c.aug.n);
    c.aug.K_inv_y = dalloc(c.aug.n);
    build_K_with_noise(gp->kernel, dim, gp->alpha, gp->lengths_sq, gp->noise, c.aug.X, c.aug.N, gp->derivs, g,
                       c.aug.K_chol);
    if (oracle_cholesky(c.aug.n, c.aug.K_chol) != 0) {
      free(c.aug.X); free(c.aug.y); free(c.aug.K_chol); free(c.aug.K_inv_y);
      goto done_early;
    }
    double* normals = dalloc((size_t)Q * num_mc);
    double* best_pts = dalloc((size_t)dim * num_mc);
    for (size_t e = 0; e < (size_t)dim * num_mc; ++e) best_pts[e] = 1.0; /* ref: .cpp:162 */
    double total = 0.0;
    for (int it = 0; it < num_mc; ++it) {
      double* z = normals + (size_t)it * Q;
      if (it % 2 == 1) {
        for (int j = 0; j < Q; ++j) z[j] = -normals[(size_t)(it - 1) * Q + j];
      } else {
        for (int j = 0; j < Q; ++j) z[j] = draw_table(src);
      }
      /* fantasy observations y_f = mu + L z written into the augmented GP, K^-1 (y - mean) re-solved
       * with the mean frozen (ref: .cpp:102-107, gpp_math.cpp:1739-1747, 531-551) */
      double* yf = c.aug.y + n;
      memcpy(yf, mu, (size_t)Q * sizeof(double));
      gemv(L, 0, z, 1.0, 1.0, Q, Q, Q, yf);
      memcpy(c.aug.K_inv_y, c.aug.y, (size_t)c.aug.n * sizeof(double));
      for (int i = 0; i < c.aug.N; ++i) c.aug.K_inv_y[(size_t)i * bs] -= c.aug.mean;
      oracle_potrs(c.aug.K_chol, c.aug.n, 1, c.aug.K_inv_y);
      double bfv = 0.0;
      optimal_posterior_mean(&c, gd, inner_bounds, dset, M, best_pts + (size_t)it * dim, &bfv);
      total += best_post + bfv;
    }
    result = total / (double)num_mc;
    if (best_points_out) memcpy(best_points_out, best_pts, (size_t)dim * num_mc * sizeof(double));

    if (grad) {
      /* ref: .cpp:134-161, 199-225 and gpp_math.cpp:788-839, 1063-1126, 1601-1651 */
      double* agg = dzero((size_t)dim * q);
      double* gm = dalloc((size_t)dim * q * bs);
      gp_grad_mean(gp, &st, gm);
      if (winner >= 0 && winner < q)
        for (int d = 0; d < dim; ++d) agg[winner * dim + d] += num_mc * gm[d + (size_t)winner * bs * dim];
      double* gchol = dalloc((size_t)dim * Q * Q);
      double* kt = dalloc((size_t)n);
      double* kinv_kt = dalloc((size_t)n);
      double* cov = dalloc(Q);
      double* w = dalloc(Q);
      double* gcv = dalloc((size_t)dim * Q);
      double* blk = dalloc((size_t)dim * bs);
      double* dLw = dalloc(Q);
      double* t1 = dalloc(Q);
      for (int k = 0; k < q; ++k) {
        gp_grad_chol_point(gp, &st, k, L, gchol);
        for (int it = 0; it < num_mc; ++it) {
          const double* xs = best_pts + (size_t)it * dim;
          const double* z = normals + (size_t)it * Q;
          /* Cov_n(union, x*) = K(union, x*) - (K^-1 K*)^T K(X, x*) */
          build_mix_cov(gp->kernel, dim, gp->alpha, gp->lengths_sq, gp->X, gp->N, gp->derivs, g, xs, 1, NULL, 0, kt);
          build_mix_cov(gp->kernel, dim, gp->alpha, gp->lengths_sq, un, U, gp->derivs, g, xs, 1, NULL, 0, cov);
          gemv(st.KinvKs, 1, kt, -1.0, 1.0, n, Q, n, cov);
          memcpy(w, cov, (size_t)Q * sizeof(double));
          oracle_trsv(L, 0, Q, Q, w); /* w = L^-1 Cov */
          /* d Cov / d Xq_k (rows of point k only) */
          memcpy(kinv_kt, kt, (size_t)n * sizeof(double));
          oracle_potrs(gp->K_chol, n, 1, kinv_kt);
          memset(gcv, 0, (size_t)dim * Q * sizeof(double));
          grad_cov_block(gp->kernel, dim, gp->alpha, gp->lengths_sq, un + (size_t)k * dim, gp->derivs, g, xs, NULL,
                         0, blk);
          for (int m = 0; m < bs; ++m) {
            const int row = k * bs + m;
            const double* gks = st.grad_K_star + (size_t)row * dim * n;
            for (int d = 0; d < dim; ++d) {
              double s = 0.0;
              for (int j = 0; j < n; ++j) s += gks[d + (size_t)j * dim] * kinv_kt[j];
              gcv[d + (size_t)row * dim] = blk[d + m * dim] - s;
            }
          }
          for (int d = 0; d < dim; ++d) {
            /* temp = L^-1 dCov - L^-1 dL w */
            for (int r = 0; r < Q; ++r) t1[r] = gcv[d + (size_t)r * dim];
            oracle_trsv(L, 0, Q, Q, t1);
            for (int r = 0; r < Q; ++r) {
              double s = 0.0;
              for (int cc = 0; cc <= r; ++cc) s += gchol[d + (size_t)cc * dim + (size_t)r * dim * Q] * w[cc];
              dLw[r] = s;
            }
            oracle_trsv(L, 0, Q, Q, dLw);
            double dot = 0.0;
            for (int r = 0; r < Q; ++r) dot += (t1[r] - dLw[r]) * z[r];
            agg[k * dim + d] -= dot;
          }
        }
      }
      for (int e = 0; e < q * dim; ++e) grad[e] = agg[e] / (double)num_mc;
      free(agg); free(gm); free(gchol); free(kt); free(kinv_kt); free(cov); free(w); free(gcv); free(blk);
      free(dLw); free(t1);
    }
    free(normals); free(best_pts);
    free(c.aug.X); free(c.aug.y); free(c.aug.K_chol); free(c.aug.K_inv_y);
  }
done_early:
  free(un); free(dset); free(mu); free(L);
  pstate_free(&st);
  return result;
}

/* ref: gpp_python_knowledge_gradient.cpp:306-350 -> ComputeOptimalPosteriorMean with one start on the GP itself */
double oracle_posterior_mean_optimization(const oracle_gp* gp, int num_fidelity, const double* gd,
                                          const double* bounds, const double* initial_guess, double* best_point) {
  kg_ctx c;
  c.base = gp;
  c.nf = num_fidelity;
  c.aug = *gp;
  double best_value = 0.0;
  optimal_posterior_mean(&c, gd, bounds, initial_guess, 1, best_point, &best_value);
  return best_value;
}

double oracle_kg(const oracle_gp* gp, int num_fidelity, const double* gd, const double* inner_bounds,
                 const double* discrete_pts, int num_pts, const double* Xq, const double* Xp, int q, int p,
                 int num_mc, double best_so_far, const double* table, int table_len, double* grad,
                 double* best_points) {
  normal_src src = {table, table_len, 0};
  return kg_impl(gp, num_fidelity, gd, inner_bounds, discrete_pts, num_pts, Xq, Xp, q, p, num_mc, best_so_far, &src,
                 grad, best_points);
}

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11) + Box-Muller: the CUDA path's normal stream, restated on the host so
 * the oracle / reference can be fed exactly the draws the GPU consumes.
 * counter = (draw_lo, draw_hi, k, 0), key = (seed_lo, seed_hi); output words (r0..r3):
 *   u1 = ((r0<<32 | r1) >> 11 + 0.5) * 2^-53,  u2 = ((r2<<32 | r3) >> 11 + 0.5) * 2^-53
 *   n0 = sqrt(-2 ln u1) cos(2 pi u2),  n1 = sqrt(-2 ln u1) sin(2 pi u2)
 * Normal index j of draw i comes from call k = j/2, element j%2.
 * ---------------------------------------------------------------------------------------------- */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

static void philox_pair(uint64_t seed, uint64_t draw, uint32_t k, double* n0, double* n1) {
  uint32_t c[4] = {(uint32_t)draw, (uint32_t)(draw >> 32), k, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t a = ((uint64_t)c[0] << 32) | c[1];
  const uint64_t b = ((uint64_t)c[2] << 32) | c[3];
  const double u1 = ((double)(a >> 11) + 0.5) * 0x1.0p-53;
  const double u2 = ((double)(b >> 11) + 0.5) * 0x1.0p-53;
  const double r = sqrt(-2.0 * log(u1));
  const double th = 6.283185307179586476925286766559 * u2;
  *n0 = r * cos(th);
  *n1 = r * sin(th);
}

void oracle_philox_normals(uint64_t seed, uint64_t first_draw, int num_draws, int per_draw, double* out) {
  for (int i = 0; i < num_draws; ++i)
    for (int k = 0; k < (per_draw + 1) / 2; ++k) {
      double a, b;
      philox_pair(seed, first_draw + (uint64_t)i, (uint32_t)k, &a, &b);
      out[(size_t)i * per_draw + 2 * k] = a;
      if (2 * k + 1 < per_draw) out[(size_t)i * per_draw + 2 * k + 1] = b;
    }
}

/* ---- candidate-list evaluators (OpenMP over candidates; each candidate replays the same Philox table) ---- */
void oracle_kg_at_point_list(const oracle_gp* gp, int num_fidelity, const double* gd, const double* inner_bounds,
                             const double* discrete_pts, int num_pts, const double* candidates, const double* Xp,
                             int num_candidates, int q, int p, int num_mc, double best_so_far, int num_threads,
                             uint64_t seed, double* values, double* grads) {
  const int Q = (q + p) * (1 + gp->g), pairs = (num_mc + 1) / 2;
  double* table = dalloc((size_t)pairs * Q);
  oracle_philox_normals(seed, 0, pairs, Q, table);
#pragma omp parallel for schedule(static) num_threads(num_threads)
  for (int c = 0; c < num_candidates; ++c) {
    values[c] = oracle_kg(gp, num_fidelity, gd, inner_bounds, discrete_pts, num_pts,
                          candidates + (size_t)c * q * gp->dim, Xp, q, p, num_mc, best_so_far, table, pairs * Q,
                          grads ? grads + (size_t)c * q * gp->dim : NULL, NULL);
  }
  free(table);
}

void oracle_ei_at_point_list(const oracle_gp* gp, const double* candidates, const double* Xp, int num_candidates,
                             int q, int p, int num_mc, double best_so_far, int num_threads, uint64_t seed,
                             double* values, double* grads) {
  const int U = q + p;
  double* table = dalloc((size_t)num_mc * U);
  oracle_philox_normals(seed, 0, num_mc, U, table);
#pragma omp parallel for schedule(static) num_threads(num_threads)
  for (int c = 0; c < num_candidates; ++c) {
    values[c] = oracle_ei(gp, candidates + (size_t)c * q * gp->dim, Xp, q, p, num_mc, best_so_far, table,
                          num_mc * U, grads ? grads + (size_t)c * q * gp->dim : NULL);
  }
  free(table);
}

int oracle_max_threads(void) { return omp_get_max_threads(); }

/* ---- MCMC-averaged acquisition over an ensemble of GPs ---------------------------------------------------------------
 * ref: KnowledgeGradientMCMCEvaluator::ComputeCost / ComputeGradCost / ComputeKnowledgeGradient /
 * ComputeGradKnowledgeGradient (gpp_knowledge_gradient_mcmc_optimization.cpp:87-180) and
 * ExpectedImprovementMCMCEvaluator (gpp_expected_improvement_mcmc_optimization.cpp:47-85).
 * discrete_pts[num_gp][num_pts][dim-nf], best_so_far[num_gp]; every member replays the same table. */
double oracle_kg_mcmc(const oracle_gp* const* gps, int num_gp, int num_fidelity, const double* gd,
                      const double* inner_bounds, const double* discrete_pts, int num_pts, const double* Xq,
                      const double* Xp, int q, int p, int num_mc, const double* best_so_far, const double* table,
                      int table_len, double* grad) {
  const int dim = gps[0]->dim, ps = dim - num_fidelity;
  double* tmp = dalloc((size_t)q * dim);
  double total = 0.0;
  if (grad) memset(grad, 0, (size_t)q * dim * sizeof(double));
  for (int m = 0; m < num_gp; ++m) {
    total += oracle_kg(gps[m], num_fidelity, gd, inner_bounds, discrete_pts + (size_t)m * num_pts * ps, num_pts, Xq, Xp,
                       q, p, num_mc, best_so_far[m], table, table_len, grad ? tmp : NULL, NULL);
    if (grad)
      for (int k = 0; k < q * dim; ++k) grad[k] += tmp[k];
  }
  double cost = 1.0;
  int index = -1;
  if (num_fidelity > 0) {
    cost = 0.0;
    for (int i = 0; i < q; ++i) {
      double point_cost = 1.0;
      for (int j = ps; j < dim; ++j) point_cost *= Xq[i * dim + j];
      if (cost < point_cost) {
        cost = point_cost;
        index = i;
      }
    }
  }
  if (grad) {
    const double kg = total / (double)num_gp;
    for (int k = 0; k < q * dim; ++k) {
      double gradcost = 0.0;
      if (index >= 0 && k / dim == index && k % dim >= ps) gradcost = cost / Xq[k];
      const double ga = grad[k] / (double)num_gp;
      grad[k] = (ga * cost - kg * gradcost) / (cost * cost);
    }
  }
  free(tmp);
  return total / ((double)num_gp * cost);
}

double oracle_ei_mcmc(const oracle_gp* const* gps, int num_gp, const double* Xq, const double* Xp, int q, int p,
                      int num_mc, const double* best_so_far, const double* table, int table_len, double* grad) {
  const int dim = gps[0]->dim;
  double* tmp = dalloc((size_t)q * dim);
  double total = 0.0;
  if (grad) memset(grad, 0, (size_t)q * dim * sizeof(double));
  for (int m = 0; m < num_gp; ++m) {
    total += oracle_ei(gps[m], Xq, Xp, q, p, num_mc, best_so_far[m], table, table_len, grad ? tmp : NULL);
    if (grad)
      for (int k = 0; k < q * dim; ++k) grad[k] += tmp[k];
  }
  if (grad)
    for (int k = 0; k < q * dim; ++k) grad[k] /= (double)num_gp;
  free(tmp);
  return total / (double)num_gp;
}

/* ---- log marginal likelihood ------------------------------------------------------------------------------------------
 * ref: LogMarginalLikelihoodEvaluator::FillLogLikelihoodState + ComputeLogLikelihood, gpp_model_selection.cpp:540-612:
 * K = K(X,X) + diag(noise by type) + 1e-6 I (:546-549); y centred by the mean of the function values (:556-566);
 *   log p = -1/2 y^T K^-1 y - sum_i log L_ii - n/2 log(2 pi).
 * The reference ignores a failed factorisation (:551-553); here a singular K returns -HUGE_VAL. */
double oracle_log_marginal_likelihood(int kernel, double alpha, const double* lengths, const double* X, const double* y,
                                      const double* noise, const int* derivs, int g, int dim, int N) {
  double* nz = dalloc(1 + g);
  for (int m = 0; m <= g; ++m) nz[m] = noise[m] + 1.0e-6;
  int lm = 0;
  oracle_gp* gp = oracle_gp_create(kernel, alpha, lengths, X, y, nz, derivs, g, dim, N, &lm);
  free(nz);
  if (!gp) return -HUGE_VAL;
  const int n = gp->n, bs = 1 + g;
  double term1 = 0.0, term2 = 0.0;
  for (int i = 0; i < n; ++i) {
    const double yc = gp->y[i] - ((i % bs == 0) ? gp->mean : 0.0);
    term1 += yc * gp->K_inv_y[i];
    term2 -= log(gp->K_chol[(size_t)i * n + i]);
  }
  const double kLog2Pi = 1.8378770664093453;
  const double out = -0.5 * term1 + term2 - 0.5 * (double)n * kLog2Pi;
  oracle_gp_destroy(gp);
  return out;
}

/* ---- gradient of the log marginal likelihood wrt the hyperparameters (alpha, l_1..l_dim, noise_0..noise_g) ---------
 * ref: LogMarginalLikelihoodEvaluator::ComputeGradLogLikelihood, gpp_model_selection.cpp:629-677:
 *   d/d theta_k = 1/2 a^T (dK/d theta_k) a - 1/2 tr(K^-1 dK/d theta_k),   a = K^-1 (y - m)
 * with the per-pair blocks of SquareExponential::HyperparameterGradCovariance (gpp_covariance.cpp:245-317) and
 * MaternNu2p5::HyperparameterGradCovariance (:461-489), assembled by BuildHyperparameterGradCovarianceMatrix
 * (gpp_model_selection.cpp:386-444; dK/d noise_m = 1 on the diagonal entries of observation type m).
 * Block layout: out[i_hyper + m*(dim+1) + n*(dim+1)*(1+g)], i_hyper in [0, dim].
 * Quirk kept: the Matern routine only fills the value-value entry, so with derivative observations every other entry of
 * its block is zero (the caller's buffer is zero-initialised and never written there), and at coincident points it
 * returns d/d alpha = 1 (not k/alpha = 1 * ... for alpha != 1) and zero length derivatives (:469-473). */
static void hyper_grad_block(int kernel, int dim, double alpha, const double* lsq, const double* p1, const double* p2,
                             const int* derivs, int g, double* out) {
  const int bs = 1 + g, hd = dim + 1;
  memset(out, 0, (size_t)hd * bs * bs * sizeof(double));
  const double r2 = weighted_sqdist(p1, p2, lsq, dim);
  if (kernel != 0) {
    if (r2 == 0.0) {
      out[0] = 1.0;
      return;
    }
    const double arg = SQRT5 * sqrt(r2);
    const double poly = arg + 5.0 / 3.0 * r2;
    const double e = exp(-arg);
    out[0] = (1.0 + poly) * e;
    for (int i = 0; i < dim; ++i) {
      const double len = sqrt(lsq[i]);
      const double dr2 = -2.0 * ((p1[i] - p2[i]) / len) * ((p1[i] - p2[i]) / len) / len;
      const double dr = 0.5 * dr2 / sqrt(r2);
      out[i + 1] = alpha * e * (5.0 / 3.0 * dr2 - poly * SQRT5 * dr);
    }
    return;
  }
  /* SquareExponential: every block entry is k * P with P in {1, u_a, -u_b, -u_a u_b + [a==b]/l_a^2}, u_a = (p2_a-p1_a)/l_a^2.
   * d/d alpha = entry / alpha;  d/d l_i = entry * D_i^2 / l_i^3 + k dP/d l_i  with  d u_a/d l_i = -2 u_a / l_a [a==i],
   * d (1/l_a^2)/d l_i = -2 / l_a^3 [a==i]. */
  const double k = alpha * exp(-0.5 * r2);
  for (int m = 0; m < bs; ++m) {
    for (int n = 0; n < bs; ++n) {
      const int a = m ? derivs[m - 1] : -1, b = n ? derivs[n - 1] : -1;
      const double ua = (a >= 0) ? (p2[a] - p1[a]) / lsq[a] : 0.0;
      const double ub = (b >= 0) ? (p2[b] - p1[b]) / lsq[b] : 0.0;
      double P;
      if (a < 0 && b < 0) P = 1.0;
      else if (b < 0) P = ua;
      else if (a < 0) P = -ub;
      else P = -ua * ub + ((a == b) ? 1.0 / lsq[a] : 0.0);
      double* o = out + (size_t)m * hd + (size_t)n * hd * bs;
      o[0] = k * P / alpha;
      for (int i = 0; i < dim; ++i) {
        const double len = sqrt(lsq[i]);
        const double D = p1[i] - p2[i];
        double dP = 0.0;
        if (a >= 0 && b < 0 && a == i) dP = -2.0 * ua / len;
        if (a < 0 && b >= 0 && b == i) dP = 2.0 * ub / len;
        if (a >= 0 && b >= 0) {
          if (a == i) dP += 2.0 * ua * ub / len;
          if (b == i) dP += 2.0 * ua * ub / len;
          if (a == b && a == i) dP += -2.0 / (lsq[a] * len);
        }
        o[i + 1] = k * P * (D / len) * (D / len) / len + k * dP;
      }
    }
  }
}

void oracle_grad_log_marginal_likelihood(int kernel, double alpha, const double* lengths, const double* X,
                                         const double* y, const double* noise, const int* derivs, int g, int dim, int N,
                                         double* grad /* [dim + 1 + 1 + g] */) {
  const int nh = dim + 1 + 1 + g;
  for (int k = 0; k < nh; ++k) grad[k] = 0.0;
  double* nz = dalloc(1 + g);
  for (int m = 0; m <= g; ++m) nz[m] = noise[m] + 1.0e-6;
  int lm = 0;
  oracle_gp* gp = oracle_gp_create(kernel, alpha, lengths, X, y, nz, derivs, g, dim, N, &lm);
  free(nz);
  if (!gp) return;
  const int n = gp->n, bs = 1 + g, hd = dim + 1;
  /* W = a a^T - K^-1 */
  double* Kinv = dzero((size_t)n * n);
  for (int i = 0; i < n; ++i) Kinv[(size_t)i * n + i] = 1.0;
  oracle_potrs(gp->K_chol, n, n, Kinv);
  double* blk = dalloc((size_t)hd * bs * bs);
  for (int i = 0; i < N; ++i) {    /* column point */
    for (int j = 0; j < N; ++j) {  /* row point */
      hyper_grad_block(kernel, dim, alpha, gp->lengths_sq, gp->X + (size_t)j * dim, gp->X + (size_t)i * dim, gp->derivs,
                       g, blk);
      for (int m = 0; m < bs; ++m) {
        for (int nn = 0; nn < bs; ++nn) {
          const int row = j * bs + m, col = i * bs + nn;
          const double W = gp->K_inv_y[row] * gp->K_inv_y[col] - Kinv[(size_t)col * n + row];
          for (int h = 0; h < hd; ++h) grad[h] += 0.5 * W * blk[h + (size_t)m * hd + (size_t)nn * hd * bs];
        }
      }
    }
  }
  for (int row = 0; row < n; ++row) {
    const double W = gp->K_inv_y[row] * gp->K_inv_y[row] - Kinv[(size_t)row * n + row];
    grad[hd + row % bs] += 0.5 * W;
  }
  free(blk);
  free(Kinv);
  oracle_gp_destroy(gp);
}
