// Live roofline denominators for the FP64 kernels: MEASURED_PEAKS.json only carries HBM and bf16 numbers, and the
// fused MC kernels are bound by the FP64 vector pipe (DFMA), the Cholesky trailing update by the FP64 tensor pipe
// (DMMA, mma.sync m8n8k4).  Both are measured here with CUDA events on the launching stream.
#include <algorithm>

#include "internal.cuh"

namespace cmoe {
namespace {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9, a4 = a0 + 4e-9, a5 = a0 + 5e-9,
         a6 = a0 + 6e-9, a7 = a0 + 7e-9;
  const double b = 0.999999, c = 1e-12;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Do the FP64 vector pipe (DFMA) and the FP64 tensor sub-pipe (DMMA) run concurrently?  Every warp interleaves equal
// amounts of both (8 x m8n8k4 = 2048 FMA and 64 DFMA x 32 lanes = 2048 FMA per iteration).
__global__ void __launch_bounds__(256) mixed_peak_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-9 + i * 1e-9;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9, m = 0.999999, k = 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fma(v[j], m, k);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace
}  // namespace cmoe

using namespace cmoe;  // NOLINT

// tflops[0] = FP64 FMA (vector pipe) TFLOP/s, tflops[1] = FP64 DMMA (tensor pipe) TFLOP/s
extern "C" int cmoe_bench_fp64_peaks(int device, double* tflops) {
  return guarded(nullptr, [&] {
    require_device(device);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 1 << 16;
    DevBuf<double> out(static_cast<size_t>(blocks) * threads);
    cudaStream_t s;
    CMOE_CUDA(cudaStreamCreate(&s));
    for (int which = 0; which < 2; ++which) {
      double best = 0.0;
      for (int rep = 0; rep < 4; ++rep) {
        EventTimer t;
        t.start(s);
        if (which == 0) {
          dfma_peak_kernel<<<blocks, threads, 0, s>>>(out.p, iters);
        } else {
          dmma_peak_kernel<<<blocks, threads, 0, s>>>(out.p, iters);
        }
        t.stop(s);
        const double sec = t.ms() * 1e-3;
        const double flops = (which == 0) ? 2.0 * 8.0 * iters * static_cast<double>(blocks) * threads
                                          : 2.0 * 256.0 * 8.0 * iters * static_cast<double>(blocks) * (threads / 32);
        if (rep > 0) best = std::max(best, flops / sec * 1e-12);
      }
      tflops[which] = best;
    }
    CMOE_CUDA(cudaGetLastError());
    cudaStreamDestroy(s);
  });
}

// tflops[0] = total FP64 TFLOP/s with DFMA and DMMA interleaved 1:1 (by flops) in every warp.  ~= the larger of the two
// peaks if the pipes share hardware, ~= their sum if they run concurrently.
extern "C" int cmoe_bench_fp64_mixed(int device, double* tflops) {
  return guarded(nullptr, [&] {
    require_device(device);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 1 << 14;
    DevBuf<double> out(static_cast<size_t>(blocks) * threads);
    cudaStream_t s;
    CMOE_CUDA(cudaStreamCreate(&s));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
      EventTimer t;
      t.start(s);
      mixed_peak_kernel<<<blocks, threads, 0, s>>>(out.p, iters);
      t.stop(s);
      const double flops = 2.0 * (2048.0 + 2048.0) * iters * static_cast<double>(blocks) * (threads / 32);
      if (rep > 0) best = std::max(best, flops / (t.ms() * 1e-3) * 1e-12);
    }
    tflops[0] = best;
    CMOE_CUDA(cudaGetLastError());
    cudaStreamDestroy(s);
  });
}
