// rng.cuh -- counter-based normal stream for the sm_100a kernels ("normal stream v1", include/pcf.h).
//
// Replaces the reference's std::mt19937 + std::normal_distribution (src/mc_eur.cpp:16-20,
// src/mc_asia.cpp:20-24, include/common.h:188-192) and boost::mt19937 + boost::normal_distribution
// (include/mvn.h:21-30): Philox4x32-10 keyed by the seed, counter = (global index, draw/2, stream),
// so 1/2/4/8 GPUs consume identical variates.
#pragma once
#include <cstdint>

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

// Box-Muller pair from one Philox block.
//   a = X1 >> 12, b = X2 >> 12;  u1 = 1 - a*2^-52 in (0,1];  2*u2 = (2b+1)*2^-52 in (0,2)
//   (z_even, z_odd) = sqrt(-2 ln u1) * (cos, sin)(pi * 2 u2)
// The two uniforms are built by bit injection into [1,2) + one exact FP64 op each (no I2F).
__device__ __forceinline__ void box_muller_pair(const uint32_t x[4], double& z_even, double& z_odd) {
  double d1 = __hiloint2double((int)(0x3FF00000u | (x[1] >> 12)), (int)((x[1] << 20) | (x[0] >> 12)));
  double d2 = __hiloint2double((int)(0x3FF00000u | (x[3] >> 12)), (int)((x[3] << 20) | (x[2] >> 12)));
  double u1 = 2.0 - d1;                                // exact
  double t2 = fma(d2, 2.0, -2.0 + 0x1p-52);            // exact: (2b+1)*2^-52
  double R = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(t2, &s, &c);
  z_even = R * c;
  z_odd = R * s;
}

__device__ __forceinline__ void normal_pair(const PhiloxKey& key, uint64_t index, uint32_t block,
                                            uint32_t stream, double& z_even, double& z_odd) {
  uint32_t x[4];
  philox4x32_10(key, (uint32_t)index, (uint32_t)(index >> 32), block, stream, x);
  box_muller_pair(x, z_even, z_odd);
}

}  // namespace pcf
