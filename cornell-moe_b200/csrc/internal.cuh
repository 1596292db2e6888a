// Internal declarations shared by the CUDA translation units of libcornell_moe_b200.so.
// Everything here is host-side C++ plumbing around hand-written sm_100a kernels; there is no CPU compute path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cmoe_b200.h"

namespace cmoe {

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
struct Error : public std::runtime_error {
  int code;
  int info;
  Error(int code_in, const std::string& msg, int info_in = 0) : std::runtime_error(msg), code(code_in), info(info_in) {}
};

void set_last_error(const std::string& msg);

#define CMOE_CUDA(call)                                                                                  \
  do {                                                                                                   \
    cudaError_t err__ = (call);                                                                          \
    if (err__ != cudaSuccess) {                                                                          \
      char buf__[512];                                                                                   \
      snprintf(buf__, sizeof(buf__), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
      throw ::cmoe::Error(CMOE_ERR_RUNTIME, buf__);                                                      \
    }                                                                                                    \
  } while (0)

#define CMOE_REQUIRE(cond, code, msg)                 \
  do {                                                \
    if (!(cond)) throw ::cmoe::Error((code), (msg)); \
  } while (0)

// Runs f(), translating exceptions into the C-ABI status codes.
template <class F>
int guarded(int* info, F&& f) {
  try {
    if (info) *info = 0;
    cudaGetLastError();  // a non-sticky error left behind by an earlier, unrelated call must not fail this one
    f();
    return CMOE_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    if (info) *info = e.info;
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return CMOE_ERR_RUNTIME;
  }
}

void require_device(int device);  // throws CMOE_ERR_NO_DEVICE / CMOE_ERR_BOUNDS

// ---------------------------------------------------------------------------------------------------------------
// device memory
// ---------------------------------------------------------------------------------------------------------------
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t count = 0;
  DevBuf() = default;
  explicit DevBuf(size_t n) { alloc(n); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), count(o.count) { o.p = nullptr; o.count = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; count = o.count; o.p = nullptr; o.count = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n) {
    release();
    count = n;
    if (n) CMOE_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), n * sizeof(T)));
  }
  void ensure(size_t n) { if (n > count) alloc(n); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    count = 0;
  }
  void upload(const T* host, size_t n, cudaStream_t s) {
    ensure(n);
    if (n) CMOE_CUDA(cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void download(T* host, size_t n, cudaStream_t s) const {
    if (n) CMOE_CUDA(cudaMemcpyAsync(host, p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
  void zero(cudaStream_t s) { if (count) CMOE_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), s)); }
};

struct EventTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  EventTimer() { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~EventTimer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  void start(cudaStream_t s) { cudaEventRecord(a, s); }
  void stop(cudaStream_t s) { cudaEventRecord(b, s); }
  float ms() { float t = 0.f; cudaEventSynchronize(b); cudaEventElapsedTime(&t, a, b); return t; }
};

// ---------------------------------------------------------------------------------------------------------------
// covariance kernel description, passed by value to every kernel
// ---------------------------------------------------------------------------------------------------------------
struct KernelSpec {
  int kernel;  // CMOE_KERNEL_*
  int dim;
  int g;       // number of derivative observations per sampled point
  double alpha;
  double lsq[CMOE_MAX_DIM];      // l_k^2
  double inv_len[CMOE_MAX_DIM];  // 1 / l_k
  int derivs[CMOE_MAX_DIM];      // observed partial-derivative indices of the training data
};

// ---------------------------------------------------------------------------------------------------------------
// the GP handle
// ---------------------------------------------------------------------------------------------------------------
}  // namespace cmoe

struct cmoe_gp {
  int device = 0;
  cmoe::KernelSpec spec{};
  int N = 0;  // points
  int n = 0;  // rows = N * (1 + g)
  double mean = 0.0;
  std::vector<double> hX, hy, hnoise, hlengths;
  cudaStream_t stream = nullptr;
  cmoe::DevBuf<double> dX;        // [N][dim]
  cmoe::DevBuf<double> dXs;       // [N][dim] scaled by 1/l (for the MC kernels)
  cmoe::DevBuf<double> dy;        // [n]
  cmoe::DevBuf<double> dnoise;    // [1+g]
  cmoe::DevBuf<double> dK;        // [n*n] column-major Cholesky factor (lower)
  cmoe::DevBuf<double> dKinvY;    // [n]
  cmoe::DevBuf<int> dFlag;        // [1] Cholesky failure index
  double fit_usec[3] = {0, 0, 0};
  uint64_t generation = 0;  // bumped by every (re)fit / append: explicit q-KG plans are bound to one generation
  // one cached q-KG plan (device workspace) so that repeated cmoe_kg_eval calls with the same configuration — every
  // step of an outer optimiser — do not re-allocate gigabytes of scratch; owned by kg.cu
  mutable struct cmoe_kg_plan* cached_plan = nullptr;
  // replicas of this GP on other devices (multi-GPU multistart inside one call); rebuilt when the generation changes
  mutable std::vector<cmoe_gp*> replicas;
  mutable uint64_t replicas_generation = 0;
  ~cmoe_gp();
};

namespace cmoe {

constexpr int kTrsmNB = 32;

// ---- gp.cu ----
void fit_gp(cmoe_gp* gp, bool mean_change);
// bit-identical copy of a fitted GP on another device (peer copies of the factor; no refit)
cmoe_gp* clone_gp_to_device(const cmoe_gp* gp, int device);
// ---- kg.cu ----
void drop_cached_plan(const cmoe_gp* gp);

// ---- linalg.cu -----------------------------------------------------------------------------------------------
// In-place blocked lower Cholesky of the n*n column-major matrix A (lda = n).  *flag (device) receives 0 or the
// failing leading-minor index (k+1), with the reference's pivot test (> 1e-16).  Enqueued on `s` (large systems also
// use an internal side stream for the look-ahead update); returns after the factorisation has completed.
void potrf_lower(double* A, int n, int* flag, cudaStream_t s);
// potrf_coop.cu: one cooperative launch per 256-column panel (n >= 1024, even); false = not applicable, A untouched
bool potrf_lower_coop(double* A, int n, int* flag, cudaStream_t s);
// X <- (L L^T)^-1 X for nrhs right-hand sides; X is n*nrhs column-major with leading dimension ldx.
void potrs_lower(const double* L, int n, double* X, int ldx, int nrhs, cudaStream_t s);
// X <- L^-1 X (trans = false) or L^-T X (trans = true).  nrhs <= 4 with n >= 1024 takes the single-launch chained
// solver (which synchronises `s`); everything else is the blocked multi-RHS kernel, asynchronous on `s`.
void trsm_lower(const double* L, int n, double* X, int ldx, int nrhs, bool trans, cudaStream_t s);
// trsv_coop.cu: one cooperative launch, owner + helper CTAs per 128-unknown block row; false = not applicable / gave up
bool trsv_coop(const double* L, int n, double* x, bool trans, cudaStream_t s);
bool trsv_coop_pair(const double* L, int n, double* x, cudaStream_t s);

// ---- posterior.cu -------------------------------------------------------------------------------------------
constexpr int kMaxQ = 96;  // largest (q+p)*(1+num_derivatives) handled by the per-set kernels

// Device-resident posterior quantities for a batch of nc point sets (num points each, gs derivative rows per point).
struct PosteriorBatch {
  int nc = 0, num = 0, gs = 0, Q = 0, nd = 0, n = 0, dim = 0;
  DevBuf<double> P;      // [nc][num][dim]   (filled by the caller before run())
  DevBuf<int> dPs;       // [gs]
  DevBuf<double> Ks, B;  // [n][nc*Q] column-major: K* and K^-1 K*
  DevBuf<double> mu;     // [nc][Q]
  DevBuf<double> var;    // [nc][Q*Q]
  DevBuf<double> chol;   // [nc][Q*Q] Cholesky of (var + diagonal term), upper triangle zeroed
  DevBuf<double> gmu;    // [nc][nd*(1+gs)][dim]
  DevBuf<double> E;      // [nc][nd*(1+gs)][Q][dim]
  DevBuf<double> gvar;   // [nc][nd][Q][Q][dim]
  DevBuf<double> gchol;  // [nc][nd][Q][Q][dim]
  DevBuf<int> fail;      // [nc]
  void configure(const cmoe_gp& gp, int nc, int num, const int* dPs_host, int gs, int nd, cudaStream_t s);
  // diag_mode: 0 none, 1 +1e-6 (EI), 2 +noise_variance[type] (KG)
  void run(const cmoe_gp& gp, int diag_mode, bool want_chol, bool want_grad_chol, cudaStream_t s);
  int first_failure(cudaStream_t s, int* which_set);
};

// ---- ei.cu ---------------------------------------------------------------------------------------------------
void upload_union_sets(PosteriorBatch& pb, const double* candidates, int nc, int q, const double* Xp, int p, int dim,
                       cudaStream_t s);
void ei_eval_batch(const cmoe_gp& gp, const double* candidates, int nc, int q, const double* Xp, int p, int num_mc,
                   double best_so_far, uint64_t seed, const double* dtable, double* ei_host, double* grad_host);

// ---- cov.cu --------------------------------------------------------------------------------------------------
// Lower triangle of K(X,X) + diag(noise by observation type), n*n column-major.
// Xs = X scaled by 1/l (used by the g == 0 fast path).
void build_covariance(const KernelSpec& spec, const double* X, const double* Xs, int N, const double* noise, double* K,
                      cudaStream_t s);
// K(X, P): rows = sampled rows (N*(1+g)), cols = P rows (num*(1+gs)); column-major, ld = n.
void build_mix_covariance(const KernelSpec& spec, const double* X, int N, const double* P, int num, const int* dPs,
                          int gs, double* out, cudaStream_t s);

// device-side Philox table for tests
void philox_normals_device(uint64_t seed, uint64_t first_draw, int num_draws, int per_draw, double* out,
                           cudaStream_t s);

// CMOE_LEGACY_LINALG=1 routes the large-N fit through the round-1 kernels (launch-per-step Cholesky, chained trsv,
// LDGSTS covariance build): an A/B switch for profiling, and a fall-back should a device mis-handle cooperative launches
bool legacy_linalg();
// CMOE_COV_TMA=1 builds K(X,X) with the TMA / DMMA kernel (tile stores through a tensor map); off by default: at d = 10
// it measures 62 us against 52 us for the LDGSTS kernel
bool cov_tma_enabled();
int set_option(const char* name, int value);

int launches_issued();        // global counter of kernel launches made by this library (host side)
void count_launch(int n = 1);

}  // namespace cmoe
