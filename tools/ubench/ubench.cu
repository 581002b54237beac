// Dev microbenchmarks for the FP64 / INT pipes of sm_100a (B200). Prints cycles per warp-instruction per SMSP.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/ubench/ubench tests/ubench/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096

// ---- dependent DFMA chains: CH independent chains per thread, operands: x = fma(x, a, b), a/b in registers
template <int CH>
__global__ void k_dfma_reg(double a, double b, double* out, long long* cyc) {
  double x[CH];
  double ar = a + threadIdx.x * 1e-9, br = b + threadIdx.x * 1e-9;  // force register operands
  for (int i = 0; i < CH; ++i) x[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = fma(x[i], ar, br);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// 3 distinct register-pair operands per DFMA: x_i = fma(y_i, z_i, x_i)
template <int CH>
__global__ void k_dfma_3reg(double a, double b, double* out, long long* cyc) {
  double x[CH], y[CH], z[CH];
  for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x + i; y[i] = a + i * 1e-9 + threadIdx.x * 1e-12; z[i] = b + i * 1e-9; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = fma(y[i], z[i], x[i]);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; ++i) s += x[i] + y[i] + z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// constant-bank operand: x = fma(x, x, c[..])
__constant__ double c_k[8] = {1e-9, 2e-9, 3e-9, 4e-9, 5e-9, 6e-9, 7e-9, 8e-9};
template <int CH>
__global__ void k_dfma_const(double* out, long long* cyc) {
  double x[CH];
  for (int i = 0; i < CH; ++i) x[i] = 1e-3 * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = fma(x[i], x[i], c_k[i & 7]);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// INT: LOP3 chains, IMAD.WIDE chains
template <int CH>
__global__ void k_lop3(uint32_t a, uint32_t* out, long long* cyc) {
  uint32_t x[CH], y = a + threadIdx.x, z = a * 3 + threadIdx.x;
  for (int i = 0; i < CH; ++i) x[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = x[i] ^ y ^ z, y += 1;  // LOP3 + IADD
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < CH; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + y;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
__global__ void k_imadwide(uint32_t a, uint32_t* out, long long* cyc) {
  uint32_t x[CH];
  for (int i = 0; i < CH; ++i) x[i] = threadIdx.x + i + a;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      uint64_t p = (uint64_t)x[i] * 0xD2511F53u;
      x[i] = (uint32_t)(p >> 32) ^ (uint32_t)p;  // IMAD.WIDE + LOP3
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < CH; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// mixed: CH DFMA chains + CI integer chains in the same loop
template <int CH, int CI>
__global__ void k_mixed(double a, double b, uint32_t ia, double* out, long long* cyc) {
  double x[CH];
  uint32_t u[CI + 1];
  for (int i = 0; i < CH; ++i) x[i] = 1e-3 * (threadIdx.x + i);
  for (int i = 0; i < CI; ++i) u[i] = threadIdx.x + i + ia;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = fma(x[i], x[i], c_k[i & 7]);
#pragma unroll
    for (int i = 0; i < CI; ++i) {
      uint64_t p = (uint64_t)u[i] * 0xD2511F53u;
      u[i] = (uint32_t)(p >> 32) ^ (uint32_t)p;
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; ++i) s += x[i];
  for (int i = 0; i < CI; ++i) s += u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// generic mixed kernel: CH const-operand DFMA chains + CI chains of integer op kind K
// K: 0 = IMAD.WIDE(+use both halves via LOP3), 1 = IMAD lo only, 2 = IMAD.HI only, 3 = LOP3 only, 4 = FFMA, 5 = SHF,
// 6 = IMAD.HI + IMAD lo as two instructions
template <int CH, int CI, int K>
__global__ void k_mix2(uint32_t ia, double* out, long long* cyc) {
  double x[CH + 1];
  uint32_t u[CI + 1];
  float f[CI + 1];
  for (int i = 0; i < CH; ++i) x[i] = 1e-3 * (threadIdx.x + i);
  for (int i = 0; i < CI; ++i) { u[i] = threadIdx.x + i + ia; f[i] = u[i] * 1e-3f; }
  uint32_t y = ia * 7 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = fma(x[i], x[i], c_k[i & 7]);
#pragma unroll
    for (int i = 0; i < CI; ++i) {
      if (K == 0) { uint64_t p = (uint64_t)u[i] * 0xD2511F53u; u[i] = (uint32_t)(p >> 32) + (uint32_t)p; }
      if (K == 1) u[i] = u[i] * 0xD2511F53u + y;
      if (K == 2) u[i] = __umulhi(u[i], 0xD2511F53u) + y;
      if (K == 3) u[i] = (u[i] ^ y) | (u[i] & 0x55555555u) + 0;
      if (K == 4) f[i] = fmaf(f[i], f[i], 1e-3f);
      if (K == 5) u[i] = __funnelshift_l(u[i], y, 7) ^ y;
      if (K == 6) {  // the same 64-bit product as K == 0 from IMAD.HI + IMAD (lo) issued separately
        uint32_t hi, lo;
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(u[i]), "r"(0xD2511F53u));
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(u[i]), "r"(0xD2511F53u));
        u[i] = hi + lo;
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; ++i) s += x[i];
  for (int i = 0; i < CI; ++i) s += u[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <class F>
void run(const char* name, int threads, int instr_per_iter_per_warp, F launch) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, sizeof(double) * 148 * 1024 * 2); cudaMalloc(&cyc, 8);
  launch(out, cyc); launch(out, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  int warps_per_smsp = threads / 32 / 4; if (warps_per_smsp < 1) warps_per_smsp = 1;
  double per_instr = (double)h / ITERS / instr_per_iter_per_warp;           // cycles per instr of ONE warp
  double per_smsp = per_instr / (threads >= 128 ? warps_per_smsp : 1);      // cycles per warp-instr per SMSP
  printf("%-44s threads=%4d cyc/iter=%8.2f  cyc/instr/warp=%6.2f  cyc/warp-instr/SMSP=%6.2f  %s\n", name, threads,
         (double)h / ITERS, per_instr, per_smsp, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
#define RUN_D(K, CH, TH) run(#K "<" #CH ">", TH, CH, [&](double* o, long long* c) { K<CH><<<1, TH>>>(0.999, 1e-9, o, c); })
  RUN_D(k_dfma_reg, 1, 32); RUN_D(k_dfma_reg, 2, 32); RUN_D(k_dfma_reg, 4, 32); RUN_D(k_dfma_reg, 8, 32);
  RUN_D(k_dfma_reg, 1, 256); RUN_D(k_dfma_reg, 2, 256); RUN_D(k_dfma_reg, 4, 256); RUN_D(k_dfma_reg, 1, 1024); RUN_D(k_dfma_reg, 4, 1024);
  RUN_D(k_dfma_3reg, 1, 32); RUN_D(k_dfma_3reg, 4, 32); RUN_D(k_dfma_3reg, 4, 256); RUN_D(k_dfma_3reg, 4, 1024); RUN_D(k_dfma_3reg, 8, 512);
#define RUN_C(CH, TH) run("k_dfma_const<" #CH ">", TH, CH, [&](double* o, long long* c) { k_dfma_const<CH><<<1, TH>>>(o, c); })
  RUN_C(1, 32); RUN_C(4, 32); RUN_C(4, 256); RUN_C(4, 1024); RUN_C(8, 512);
#define RUN_I(K, CH, TH, PER) run(#K "<" #CH ">", TH, CH * PER, [&](double* o, long long* c) { K<CH><<<1, TH>>>(12345u, (uint32_t*)o, c); })
  RUN_I(k_lop3, 1, 32, 2); RUN_I(k_lop3, 4, 32, 2); RUN_I(k_lop3, 4, 256, 2); RUN_I(k_lop3, 4, 1024, 2);
  RUN_I(k_imadwide, 1, 32, 2); RUN_I(k_imadwide, 2, 32, 2); RUN_I(k_imadwide, 4, 32, 2); RUN_I(k_imadwide, 4, 256, 2); RUN_I(k_imadwide, 4, 1024, 2);
#define RUN_M(CH, CI, TH) run("k_mixed<" #CH "," #CI ">", TH, CH + 2 * CI, [&](double* o, long long* c) { k_mixed<CH, CI><<<1, TH>>>(0.999, 1e-9, 7u, o, c); })
  RUN_M(4, 1, 256); RUN_M(4, 2, 256); RUN_M(4, 4, 256); RUN_M(4, 2, 1024); RUN_M(4, 4, 1024); RUN_M(2, 2, 1024); RUN_M(2, 1, 1024);
#define RUN_X(CH, CI, K, TH) run("k_mix2<dfma=" #CH ",int=" #CI ",kind=" #K ">", TH, 1, [&](double* o, long long* c) { k_mix2<CH, CI, K><<<1, TH>>>(7u, o, c); })
  printf("--- k_mix2: cyc/iter column = cycles per loop iteration per 8 warps/SMSP (divide by 8 for per-warp)\n");
  RUN_X(0, 4, 0, 1024); RUN_X(0, 4, 1, 1024); RUN_X(0, 4, 2, 1024); RUN_X(0, 4, 3, 1024); RUN_X(0, 4, 4, 1024); RUN_X(0, 4, 5, 1024);
  RUN_X(4, 0, 0, 1024);
  RUN_X(4, 4, 0, 1024); RUN_X(4, 4, 1, 1024); RUN_X(4, 4, 2, 1024); RUN_X(4, 4, 3, 1024); RUN_X(4, 4, 4, 1024); RUN_X(4, 4, 5, 1024);
  RUN_X(4, 8, 3, 1024); RUN_X(4, 8, 4, 1024); RUN_X(4, 2, 0, 1024);
  RUN_X(0, 4, 6, 1024); RUN_X(4, 4, 6, 1024); RUN_X(4, 2, 6, 1024); RUN_X(8, 2, 0, 1024); RUN_X(8, 2, 6, 1024); RUN_X(8, 2, 2, 1024);
  return 0;
}
