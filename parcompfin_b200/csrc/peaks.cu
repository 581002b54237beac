// peaks.cu -- roofline denominators measured on the device the job runs on.
//   * FP64: MEASURED_PEAKS.json has no FP64 figure, so the DFMA issue peak is
//     measured here: 8 independent register-resident FMA chains per thread, full occupancy.
//   * HBM: plain device copy (read + write bytes), the same definition MEASURED_PEAKS.json uses.
#include "common.cuh"

namespace pcf {

constexpr int kPeakBlock = 256;
constexpr int kChains = 8;
constexpr int kInner = 512;

// The addend is a constant-bank operand (the form every Horner step of the production kernels uses): DFMA
// then issues at its true rate of one warp instruction per 2 cycles per SM sub-partition; with three
// register operands the register file limits it to ~2.2-2.5 cycles (tools/ubench, profiles/r1_ubench_pipes.log).
__constant__ double c_peak_addend[8] = {1e-9, 2e-9, 3e-9, 4e-9, 5e-9, 6e-9, 7e-9, 8e-9};

__global__ void __launch_bounds__(kPeakBlock) dfma_chain_kernel(double a, double b, int outer, double* sink) {
  double x[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = (double)(threadIdx.x + i) * 1e-3 * a + b;
  for (int o = 0; o < outer; ++o) {
#pragma unroll 8
    for (int it = 0; it < kInner; ++it) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) x[i] = fma(x[i], x[i], c_peak_addend[i]);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += x[i];
  if (s == 123.456) sink[0] = s;  // keeps the chains alive, never true in practice
}

__global__ void __launch_bounds__(kPeakBlock) copy_kernel(const double2* __restrict__ src,
                                                          double2* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

int run_fp64_peak(Ctx& c, double seconds_target, double* dfma_per_sec) {
  const int grid = c.sm_count * 8;
  int outer = 8;
  double best = 0.0;
  // calibrate, then one long launch (sustained clocks) -- best of the long launches
  for (int rep = 0; rep < 4; ++rep) {
    PCF_CUDA(cudaEventRecord(c.ev0, c.stream));
    dfma_chain_kernel<<<grid, kPeakBlock, 0, c.stream>>>(0.999999, 1e-9, outer, c.d_out + 40);
    PCF_CUDA(cudaGetLastError());
    PCF_CUDA(cudaEventRecord(c.ev1, c.stream));
    PCF_CUDA(cudaStreamSynchronize(c.stream));
    float ms = 0.f;
    PCF_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    double n = (double)grid * kPeakBlock * (double)outer * kInner * kChains;
    double rate = n / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
    double want = seconds_target > 0 ? seconds_target : 0.2;
    double scale = want / (ms * 1e-3 + 1e-9);
    if (scale > 64) scale = 64;
    if (rep == 0) outer = (int)(outer * scale) + 1;
  }
  *dfma_per_sec = best;
  return PCF_OK;
}

int run_hbm_peak(Ctx& c, long long bytes, double* bytes_per_sec) {
  if (bytes < (1 << 20)) bytes = 1 << 20;
  bytes &= ~255LL;
  void *a = nullptr, *b = nullptr;
  PCF_CUDA(cudaMalloc(&a, (size_t)bytes));
  if (cudaMalloc(&b, (size_t)bytes) != cudaSuccess) { cudaFree(a); cudaGetLastError(); return PCF_ENOMEM; }
  cudaMemsetAsync(a, 1, (size_t)bytes, c.stream);
  cudaMemsetAsync(b, 2, (size_t)bytes, c.stream);
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(c.ev0, c.stream);
    copy_kernel<<<c.sm_count * 16, kPeakBlock, 0, c.stream>>>((const double2*)a, (double2*)b, bytes / 16);
    cudaEventRecord(c.ev1, c.stream);
    cudaStreamSynchronize(c.stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c.ev0, c.ev1);
    double rate = 2.0 * (double)bytes / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaFree(a);
  cudaFree(b);
  PCF_CUDA(cudaGetLastError());
  *bytes_per_sec = best;
  return PCF_OK;
}

}  // namespace pcf
