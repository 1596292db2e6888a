// Latency microbenchmarks behind the design of the Cholesky pivot chain (one warp, dependent operations): how many
// cycles do a dependent DFMA, the RSQ64H seed, a shuffle, a shared-memory round trip and a block fence cost on this
// part, and how long do the 32x32 warp factorisations take per column.  Reported in cycles (clock64), one CTA.
#include "device_math.cuh"
#include "internal.cuh"
#include "linalg_dev.cuh"

namespace cmoe {
namespace {

template <int WHAT>
__global__ void __launch_bounds__(32) chain_kernel(double* out, long long* cycles, int iters) {
  __shared__ double sm[128];
  __shared__ volatile int flag;
  const int lane = threadIdx.x;
  double x = 1.0 + lane * 1e-3, acc = 0.5;
  sm[lane] = x;
  sm[lane + 32] = 0.0;
  __syncwarp();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (WHAT == 0) {  // dependent DFMA
      x = fma(x, 0.999999, 1e-9);
    } else if (WHAT == 1) {  // RSQ64H seed + one dependent multiply
      double y;
      asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
      x = y * 1.000001 + 0.5;
    } else if (WHAT == 2) {  // shuffle
      x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31) + 1e-9;
    } else if (WHAT == 3) {  // STS -> __syncwarp -> LDS (other lane's value)
      sm[lane] = x;
      __syncwarp();
      x = sm[(lane + 1) & 31] + 1e-9;
      __syncwarp();
    } else if (WHAT == 4) {  // block fence by one lane + reconvergence
      if (lane == 0) {
        __threadfence_block();
        flag = i;
      }
      x = __shfl_sync(0xffffffffu, x, 0) + 1e-9;
    } else if (WHAT == 5) {  // library rsqrt
      x = rsqrt(x) + 0.5;
    } else if (WHAT == 6) {  // library sqrt + divide
      x = 1.0 / sqrt(x) + 0.5;
    }
    acc += x;
  }
  const long long t1 = clock64();
  out[lane] = acc;
  if (lane == 0) cycles[0] = t1 - t0;
}

// 32 x 32 factorisations on a resident SPD block: VARIANT 0 = chol32_warp<true>, 1 = chol32_warp_pair
template <int VARIANT>
__global__ void __launch_bounds__(32) chol32_bench_kernel(const double* A, double* out, long long* cycles, int reps) {
  __shared__ double LT[NB * LTS + NB];
  __shared__ double colbuf[256];
  __shared__ double rd[NB];
  const int lane = threadIdx.x;
  for (int e = lane; e < NB * LTS + NB; e += 32) LT[e] = 0.0;
  for (int e = lane; e < 256; e += 32) colbuf[e] = 0.0;
  __syncwarp();
  double sink = 0.0;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    double a[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? A[c * 32 + lane] : 0.0;
    int f;
    if (VARIANT == 0) {
      f = chol32_warp<true>(a, lane, colbuf, LT, rd, 0);
    } else if (VARIANT == 1) {
      f = chol32_warp_pair<false>(a, lane, colbuf, LT, rd, 0);
    } else if (VARIANT == 2) {
      f = chol32_warp_pair<true>(a, lane, colbuf, LT, rd, 0);
    } else {
      f = chol32_warp_pipe(a, lane, colbuf, LT, rd, 0);
    }
    sink += LT[31 * LTS + 31] + f;
  }
  const long long t1 = clock64();
  out[lane] = sink;
  if (lane == 0) cycles[0] = t1 - t0;
}


// The same factorisation inside a 512-thread CTA compiled under the cooperative kernel's register cap (128):
// MODE 0: warp 0 factors, the other warps wait at the barrier; MODE 1: warps 1, 2 trail it as followers (solve_steps_rot
// paced by the progress counter, as in potrf_coop.cu).
template <int MODE>
__global__ void __launch_bounds__(512, 1) chol32_cta_kernel(const double* A, double* out, long long* cycles, int reps) {
  __shared__ double LT[NB * LTS + NB];
  __shared__ double colbuf[256];
  __shared__ double rd[NB];
  __shared__ double sink_s[64];
  __shared__ int prog_s;
  volatile int* prog = &prog_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < NB * LTS + NB; e += 512) LT[e] = 0.0;
  for (int e = tid; e < 256; e += 512) colbuf[e] = 0.0;
  __syncthreads();
  double sink = 0.0;
  long long t0 = 0, total = 0;
  for (int r = 0; r < reps; ++r) {
    if (tid == 0) *prog = 0;
    __syncthreads();
    if (tid == 0) t0 = clock64();
    if (warp == 0) {
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? A[c * 32 + lane] : 0.0;
      const int f = chol32_warp_pipe(a, lane, colbuf, LT, rd, 0, prog);
      sink += f;
    } else if (MODE == 1 && warp <= 2) {
      double x[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) x[c] = (c == lane) ? 1.0 : 0.0;
#pragma unroll 1
      for (int kb = 0; kb < 32; kb += 4) {
        while (*prog < kb + 4) __nanosleep(64);
        solve_steps_rot<32, 4, 4, true>(x, LT, rd, kb, [&](int k, double v) { sink_s[(warp - 1) * 32 + lane] = v + k; });
      }
    }
    __syncthreads();
    if (tid == 0) total += clock64() - t0;
  }
  out[tid & 31] = sink + sink_s[lane];
  if (tid == 0) cycles[0] = total;
}

}  // namespace
}  // namespace cmoe

using namespace cmoe;  // NOLINT

// out[11]; out[0..6]: cycles per iteration of the dependent chains (DFMA, RSQ64H+1, shuffle, smem round trip, block fence,
// library rsqrt, sqrt+divide); out[7..9]: cycles per COLUMN of chol32_warp<true>, chol32_warp_pair<false> and chol32_warp_pair<true> (lean pivots)
extern "C" int cmoe_bench_chain_latencies(int device, double* out) {
  return guarded(nullptr, [&] {
    require_device(device);
    DevBuf<double> sink(64), dA(32 * 32);
    DevBuf<long long> cyc(1);
    cudaStream_t s;
    CMOE_CUDA(cudaStreamCreate(&s));
    std::vector<double> hA(32 * 32);
    for (int c = 0; c < 32; ++c)
      for (int r = 0; r < 32; ++r) hA[c * 32 + r] = (r == c ? 40.0 : 0.0) + 1.0 / (1.0 + std::abs(r - c));
    dA.upload(hA.data(), hA.size(), s);
    const int iters = 4096;
    auto run = [&](int which) {
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (which) {
          case 0: chain_kernel<0><<<1, 32, 0, s>>>(sink.p, cyc.p, iters); break;
          case 1: chain_kernel<1><<<1, 32, 0, s>>>(sink.p, cyc.p, iters); break;
          case 2: chain_kernel<2><<<1, 32, 0, s>>>(sink.p, cyc.p, iters); break;
          case 3: chain_kernel<3><<<1, 32, 0, s>>>(sink.p, cyc.p, iters); break;
          case 4: chain_kernel<4><<<1, 32, 0, s>>>(sink.p, cyc.p, iters); break;
          case 5: chain_kernel<5><<<1, 32, 0, s>>>(sink.p, cyc.p, iters); break;
          case 6: chain_kernel<6><<<1, 32, 0, s>>>(sink.p, cyc.p, iters); break;
          case 7: chol32_bench_kernel<0><<<1, 32, 0, s>>>(dA.p, sink.p, cyc.p, 64); break;
          case 8: chol32_bench_kernel<1><<<1, 32, 0, s>>>(dA.p, sink.p, cyc.p, 64); break;
          case 9: chol32_bench_kernel<2><<<1, 32, 0, s>>>(dA.p, sink.p, cyc.p, 64); break;
          case 10: chol32_bench_kernel<3><<<1, 32, 0, s>>>(dA.p, sink.p, cyc.p, 64); break;
          case 11: chol32_cta_kernel<0><<<1, 512, 0, s>>>(dA.p, sink.p, cyc.p, 64); break;
          default: chol32_cta_kernel<1><<<1, 512, 0, s>>>(dA.p, sink.p, cyc.p, 64); break;
        }
        CMOE_CUDA(cudaMemcpyAsync(&c, cyc.p, sizeof(long long), cudaMemcpyDeviceToHost, s));
        CMOE_CUDA(cudaStreamSynchronize(s));
      }
      return static_cast<double>(c);
    };
    for (int w = 0; w < 7; ++w) out[w] = run(w) / iters;
    out[7] = run(7) / (64.0 * 32.0);
    out[8] = run(8) / (64.0 * 32.0);
    out[9] = run(9) / (64.0 * 32.0);
    out[10] = run(10) / (64.0 * 32.0);
    out[11] = run(11) / (64.0 * 32.0);
    out[12] = run(12) / (64.0 * 32.0);
    CMOE_CUDA(cudaGetLastError());
    cudaStreamDestroy(s);
  });
}
