// rng.cuh -- counter-based normal stream for the sm_100a kernels ("normal stream v1", include/pcf.h).
//
// Replaces the reference's std::mt19937 + std::normal_distribution (src/mc_eur.cpp:16-20,
// src/mc_asia.cpp:20-24, include/common.h:188-192) and boost::mt19937 + boost::normal_distribution
// (include/mvn.h:21-30): Philox4x32-10 keyed by the seed, counter = (global index, draw/2, stream),
// so 1/2/4/8 GPUs consume identical variates.
#pragma once
#include <cstdint>
#include "fastmath.cuh"

namespace pcf {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

struct PhiloxKey {
  // the ten round keys are loop-invariant: precomputed once per thread, they live in registers or
  // are rematerialised as immediates by ptxas when the seed is a kernel constant
  uint32_t k0[10], k1[10];
  __device__ __forceinline__ explicit PhiloxKey(uint64_t seed) {
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      k0[i] = a;
      k1[i] = b;
      a += kPhiloxW0;
      b += kPhiloxW1;
    }
  }
};

// One Philox4x32-10 block: 10 rounds x (2 IMAD.WIDE.U32 + 2 LOP3) = 40 integer instructions.
__device__ __forceinline__ void philox4x32_10(const PhiloxKey& key, uint32_t c0, uint32_t c1,
                                              uint32_t c2, uint32_t c3, uint32_t out[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)kPhiloxM0 * c0;  // mul.wide.u32
    uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.k0[i];
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k1[i];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// Box-Muller pair from one Philox block ("normal stream v1", include/pcf.h):
//   X1 = x1:x0, X2 = x3:x2;  a = X1 >> 12;  u1 = 1 - a 2^-52 in (0,1];  u2 = ((X2 >> 6) + 1/2) 2^-58 in (0,1)
//   (z_even, z_odd) = sqrt(-2 ln u1) * (cos, sin)(2 pi u2)
// u1 is built by bit injection into [1,2) and one exact subtraction (no I2F); the angle never exists as
// a double: its top 6 bits pick a sector, the other 52 fill a mantissa (fastmath.cuh).
// 33 FP64-pipe instructions per pair (CUDA math library: 67).
__device__ __forceinline__ void box_muller_pair(const uint32_t x[4], const TableView& tv, const Hoisted& hc,
                                                double& z_even, double& z_odd) {
  const double d1 = __hiloint2double((int)(0x3FF00000u | (x[1] >> 12)), (int)((x[1] << 20) | (x[0] >> 12)));
  const double u1 = 2.0 - d1;  // exact
  const double R = sqrt_pos(neg2log_unit(u1, tv, hc));
  double c, s;
  sincos_2pi_bits(x[2], x[3], tv, hc, c, s);
  z_even = R * c;
  z_odd = R * s;
}

__device__ __forceinline__ void normal_pair(const PhiloxKey& key, uint64_t index, uint32_t block,
                                            uint32_t stream, const TableView& tv, const Hoisted& hc,
                                            double& z_even, double& z_odd) {
  uint32_t x[4];
  philox4x32_10(key, (uint32_t)index, (uint32_t)(index >> 32), block, stream, x);
  box_muller_pair(x, tv, hc, z_even, z_odd);
}

// Stages the lookup tables of fastmath.cuh from global memory into this block's dynamic shared memory,
// replicated per bank group, and returns the calling thread's view. Ends with __syncthreads().
__device__ __forceinline__ TableView stage_tables(const MathTables* __restrict__ g, unsigned char* smem) {
  Pair* ln_s = reinterpret_cast<Pair*>(smem);
  Pair* sc_s = ln_s + kLnEntries * kRep16;
  double* ex_s = reinterpret_cast<double*>(sc_s + kScEntries * kRep16);
  for (int i = threadIdx.x; i < kLnEntries * kRep16; i += blockDim.x) ln_s[i] = g->ln_tab[i / kRep16];
  for (int i = threadIdx.x; i < kScEntries * kRep16; i += blockDim.x) sc_s[i] = g->sc_tab[i / kRep16];
  for (int i = threadIdx.x; i < kExpEntries * kRep8; i += blockDim.x) ex_s[i] = g->exp_tab[i / kRep8];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  TableView tv;
  tv.ln_tab = ln_s + (lane & (kRep16 - 1));
  tv.sc_tab = sc_s + (lane & (kRep16 - 1));
  tv.exp_tab = ex_s + (lane & (kRep8 - 1));
  tv.stride16 = kRep16;
  tv.stride8 = kRep8;
  return tv;
}

// |x| bound under which exp_small() may replace exp_table(): |x| <= |a| + |b| * kZMax with
// kZMax = sqrt(2 * 52 ln 2), the largest |z| the stream can produce (u1 >= 2^-52).
constexpr double kZMax = 8.5;
constexpr double kSmallExpBound = 0.11;

template <bool kSmall>
__device__ __forceinline__ double exp_any(double x, const TableView& tv, const Hoisted& hc) {
  return kSmall ? exp_small(x, hc) : exp_table(x, tv);
}

}  // namespace pcf
