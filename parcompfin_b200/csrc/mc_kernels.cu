// mc_kernels.cu -- European (a1) and Asian (a3) Monte Carlo kernels, sm_100a FP64 (the basket: basket_kernels.cu).
// One path per thread iteration, state in registers, Philox normals generated in-register,
// compensated (sum, sum^2) reduced thread -> warp -> block -> grid. No tensor cores: nothing here
// is a dense contraction; the bound is the FP64 pipe.
#include "common.cuh"
#include "reduce.cuh"
#include "rng.cuh"

namespace pcf {

constexpr int kBlock = 256;
constexpr int kBlocksPerSM = 4;  // resident CTAs per SM the grids are sized for (registers: <= 64/thread)

// ------------------------------------------------------------------------------------------------
// a1  reference src/mc_eur.cpp:23-26
struct EurArgs {
  double S0, E, drift, sigma, sqrtT;  // drift = (r - sigma^2/2) T
  int cp;
  long long N;        // global path count
  long long k0, k1;   // this GPU's pair range: paths 2k, 2k+1
  unsigned long long seed;
  const double* w;    // replay: w[n - 2*k0], already N(0,T)
};

// kPairs Philox blocks (2*kPairs paths) per thread iteration: independent integer and FP64 instruction streams
// in one loop body (see mc_asia_kernel); sums are folded into the compensated totals once per iteration.
template <bool kReplay, bool kSmallExp, int kPairs, int kMinBlocks>
__global__ void __launch_bounds__(kBlock, kMinBlocks) mc_eur_kernel(EurArgs a, const MathTables* __restrict__ tables, PeerLink link,
                                                           double* partials, unsigned int* ticket, double* out) {
  __shared__ double smem[2 * 2 * 32];
  extern __shared__ __align__(16) unsigned char tab_smem[];
  const TableView tv = stage_tables(tables, tab_smem);
  Hoisted hc;
  hc.load();
  const PhiloxKey key(a.seed);
  Comp s1, s2;
  const long long T = (long long)gridDim.x * blockDim.x;
  const double S0 = a.S0, sig = kReplay ? a.sigma : a.sigma * a.sqrtT, drift = a.drift;
  const double sgn = (double)a.cp, nE = -sgn * a.E;  // cp*(S - E) = fma(sgn, S, nE): the same rounding as (double)cp*(S - E)
  for (long long base = a.k0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; base < a.k1; base += T * kPairs) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int q = 0; q < kPairs; ++q) {
      const long long k = base + q * T;
      const bool has1 = k < a.k1, has2 = has1 && (2 * k + 1 < a.N);
      double w0, w1;
      if (kReplay) {
        w0 = has1 ? a.w[2 * (k - a.k0)] : 0.0;
        w1 = has2 ? a.w[2 * (k - a.k0) + 1] : 0.0;
      } else {
        normal_pair(key, (uint64_t)k, 0u, PCF_STREAM_EUR, tv, hc, w0, w1);  // sqrt(T) folded into `sig`
      }
      // payoff (mc_eur.cpp:24) carried DOUBLED: 2 max(t, 0) = t + |t| is one DADD (|.| is an operand modifier) where
      // max costs a DSETP and two FSELs; the sums are scaled back by 1/2 and 1/4 below -- exact, so bit-identical
      const double u0 = fma(sgn, S0 * exp_any<kSmallExp>(fma(sig, w0, drift), tv, hc), nE);
      const double u1 = fma(sgn, S0 * exp_any<kSmallExp>(fma(sig, w1, drift), tv, hc), nE);
      double v0 = u0 + fabs(u0), v1 = u1 + fabs(u1);
      v0 = has1 ? v0 : 0.0;
      v1 = has2 ? v1 : 0.0;
      t1 += v0 + v1;
      t2 = fma(v0, v0, fma(v1, v1, t2));
    }
    s1.add(t1);
    s2.add(t2);
  }
  s1.hi *= 0.5;  s1.lo *= 0.5;
  s2.hi *= 0.25; s2.lo *= 0.25;
  Comp v[2] = {s1, s2};
  grid_reduce<2>(v, smem, partials, ticket, out, &link);
}

int run_mc_eur(Ctx& c, const pcf_params& p, Shard pairs, const double* d_replay, const PeerLink& link) {
  EurArgs a;
  a.S0 = p.S0; a.E = p.E; a.sigma = p.sigma; a.cp = p.cp; a.N = p.N;
  a.drift = (p.r - p.sigma * p.sigma / 2) * p.T;
  a.sqrtT = sqrt(p.T);
  a.k0 = pairs.begin; a.k1 = pairs.end;
  a.seed = p.seed; a.w = d_replay;
  const bool small = fabs(a.drift) + fabs(a.sigma * a.sqrtT) * kZMax <= kSmallExpBound;
  if (d_replay) {
    int grid = grid_for(c, pairs.size(), kBlock, 2);
    mc_eur_kernel<true, false, 1, 2><<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, c.d_ticket, final_out(c));
  } else {
    const char* v = tuning_env("PCF_EUR_VARIANT");  // <pairs per thread><CTAs per SM> (PCF_TUNING builds)
    const int variant = v ? atoi(v) : 81;
#define PCF_EUR_CASE(P, B)                                                                                        \
  case P * 10 + B: {                                                                                              \
    int grid = grid_for(c, (pairs.size() + P - 1) / P, kBlock, B);                                                \
    if (small)                                                                                                    \
      mc_eur_kernel<false, true, P, B><<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, c.d_ticket, final_out(c)); \
    else                                                                                                          \
      mc_eur_kernel<false, false, P, B><<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, c.d_ticket, final_out(c)); \
  } break;
    switch (variant) {
#ifdef PCF_TUNING
      PCF_EUR_CASE(1, 4)
      PCF_EUR_CASE(2, 2)
      PCF_EUR_CASE(2, 3)
      PCF_EUR_CASE(3, 2)
      PCF_EUR_CASE(3, 3)
      PCF_EUR_CASE(4, 1)
      PCF_EUR_CASE(4, 2)
      PCF_EUR_CASE(5, 1)
      PCF_EUR_CASE(6, 1)
#endif
      PCF_EUR_CASE(8, 1)
      default:
        set_last_error("unknown PCF_EUR_VARIANT");
        return PCF_EINVAL;
    }
#undef PCF_EUR_CASE
  }
  c.launches++;
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

// ------------------------------------------------------------------------------------------------
// a3  reference src/mc_asia.cpp:27-37
struct AsiaArgs {
  double S0, E;
  double c0;     // 1 + r dt / 2
  double ch;     // (sigma/2) * sd      (native)  |  sigma/2 (replay)
  double cs;     // sigma * sd          (native)  |  sigma   (replay)
  double adt;    // (r - sigma^2/2) dt
  double invM;   // payoff on I / M  (division kept: see kernel)
  int cp, M;
  long long n0, n1;
  unsigned long long seed;
  const double* dB;  // replay: dB[(n-n0)*M + m]
};

template <bool kSmallExp>
__device__ __forceinline__ void asia_step(double& S, double& I, double z, double ch, double c0, double cs,
                                          double adt, const TableView& tv, const Hoisted& hc) {
  I = fma(S, fma(ch, z, c0), I);                        // I += St*(1 + r dt/2 + sigma dB/2), pre-update St (:33)
  S *= exp_any<kSmallExp>(fma(cs, z, adt), tv, hc);     // St *= exp((r - sigma^2/2) dt + sigma dB)          (:34)
}

// Replay flavour: normals come from HBM in the reference's draw order (parity path, not the fast path).
__global__ void __launch_bounds__(kBlock) mc_asia_replay_kernel(AsiaArgs a, const MathTables* __restrict__ tables,
                                                                PeerLink link, double* partials,
                                                                unsigned int* ticket, double* out) {
  __shared__ double smem[2 * 2 * 32];
  extern __shared__ __align__(16) unsigned char tab_smem[];
  const TableView tv = stage_tables(tables, tab_smem);
  Hoisted hc;
  hc.load();
  Comp s1, s2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const double Md = (double)a.M;
  for (long long n = a.n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; n < a.n1; n += stride) {
    double S = a.S0, I = 0.0;
    const double* z = a.dB + (n - a.n0) * (long long)a.M;
    for (int m = 0; m < a.M; ++m) asia_step<false>(S, I, z[m], a.ch, a.c0, a.cs, a.adt, tv, hc);
    double v = payoff(I / Md, a.E, a.cp);  // :36
    s1.add(v);
    s2.add(v * v);
  }
  Comp v[2] = {s1, s2};
  grid_reduce<2>(v, smem, partials, ticket, out, &link);
}

// Native flavour. kPaths independent paths per thread are advanced in lock step: their Philox rounds and
// FP64 chains are independent instruction streams in one loop body, which gives every warp integer AND
// FP64 work to issue at any time (ILP instead of relying on 8+ resident warps being out of phase).
template <bool kSmallExp, int kPaths, int kMinBlocks>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
mc_asia_kernel(AsiaArgs a, const MathTables* __restrict__ tables, PeerLink link, double* partials,
               unsigned int* ticket, double* out) {
  __shared__ double smem[2 * 2 * 32];
  extern __shared__ __align__(16) unsigned char tab_smem[];
  const TableView tv = stage_tables(tables, tab_smem);
  Hoisted hc;
  hc.load();
  const PhiloxKey key(a.seed);
  Comp s1, s2;
  const long long T = (long long)gridDim.x * blockDim.x;
  const double Md = (double)a.M;
  const double ch = a.ch, c0 = a.c0, cs = a.cs, adt = a.adt;
  for (long long base = a.n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; base < a.n1; base += T * kPaths) {
    double S[kPaths], I[kPaths];
    uint32_t lo[kPaths], hi[kPaths];
#pragma unroll
    for (int p = 0; p < kPaths; ++p) {
      const long long n = base + p * T;  // may run past n1 in the last sweep: computed, not accumulated
      S[p] = a.S0;
      I[p] = 0.0;
      lo[p] = (uint32_t)n;
      hi[p] = (uint32_t)((uint64_t)n >> 32);
    }
    int m = 0;
    // (unrolling this loop 2x / 3x so that consecutive date pairs share a basic block is slower at 6 paths per thread --
    // 706 -> 778 / 821 ms, the body already needs 250 registers -- and 4 paths x 2 only ties: profiles/r1s_tune_asia_unroll.log)
    for (; m + 1 < a.M; m += 2) {
#pragma unroll
      for (int p = 0; p < kPaths; ++p) {
        uint32_t x[4];
        philox4x32_10(key, lo[p], hi[p], (uint32_t)(m >> 1), PCF_STREAM_ASIA, x);
        double z0, z1;
        box_muller_pair(x, tv, hc, z0, z1);
        asia_step<kSmallExp>(S[p], I[p], z0, ch, c0, cs, adt, tv, hc);
        asia_step<kSmallExp>(S[p], I[p], z1, ch, c0, cs, adt, tv, hc);
      }
    }
    if (m < a.M) {
#pragma unroll
      for (int p = 0; p < kPaths; ++p) {
        uint32_t x[4];
        philox4x32_10(key, lo[p], hi[p], (uint32_t)(m >> 1), PCF_STREAM_ASIA, x);
        double z0, z1;
        box_muller_pair(x, tv, hc, z0, z1);
        asia_step<kSmallExp>(S[p], I[p], z0, ch, c0, cs, adt, tv, hc);
      }
    }
#pragma unroll
    for (int p = 0; p < kPaths; ++p) {
      if (base + p * T < a.n1) {
        double v = payoff(I[p] / Md, a.E, a.cp);  // :36
        s1.add(v);
        s2.add(v * v);
      }
    }
  }
  Comp v[2] = {s1, s2};
  grid_reduce<2>(v, smem, partials, ticket, out, &link);
}

template <bool kSmallExp, int kPaths, int kMinBlocks>
static void launch_asia(Ctx& c, const AsiaArgs& a, long long paths, const PeerLink& link) {
  int grid = grid_for(c, (paths + kPaths - 1) / kPaths, kBlock, kMinBlocks);
  mc_asia_kernel<kSmallExp, kPaths, kMinBlocks><<<grid, kBlock, kTableSmemBytes, c.stream>>>(
      a, c.d_tables, link, c.d_partials, c.d_ticket, final_out(c));
}

int run_mc_asia(Ctx& c, const pcf_params& p, Shard paths, const double* d_replay, const PeerLink& link) {
  AsiaArgs a;
  const double dt = (double)p.T / (double)p.M;
  const double sd = sqrt(dt);
  a.S0 = p.S0; a.E = p.E; a.cp = p.cp; a.M = p.M;
  a.c0 = 1 + p.r * dt / 2;
  a.adt = (p.r - p.sigma * p.sigma / 2) * dt;
  a.ch = d_replay ? p.sigma / 2 : (p.sigma / 2) * sd;
  a.cs = d_replay ? p.sigma : p.sigma * sd;
  a.invM = 1.0 / (double)p.M;
  a.n0 = paths.begin; a.n1 = paths.end;
  a.seed = p.seed; a.dB = d_replay;
  const bool small = fabs(a.adt) + fabs(a.cs) * kZMax <= kSmallExpBound;
  if (d_replay) {
    int grid = grid_for(c, paths.size(), kBlock, kBlocksPerSM);
    mc_asia_replay_kernel<<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, c.d_ticket, final_out(c));
  } else {
    // launch shape: PCF_ASIA_VARIANT = <paths per thread><min blocks per SM>, e.g. "14", "23" (PCF_TUNING builds)
    const char* v = tuning_env("PCF_ASIA_VARIANT");
    const int variant = v ? atoi(v) : 61;
#define PCF_ASIA_CASE(P, B)                                   \
  case P * 10 + B:                                            \
    if (small) launch_asia<true, P, B>(c, a, paths.size(), link);   \
    else launch_asia<false, P, B>(c, a, paths.size(), link);        \
    break;
    switch (variant) {
#ifdef PCF_TUNING
      PCF_ASIA_CASE(1, 3)
      PCF_ASIA_CASE(1, 4)
      PCF_ASIA_CASE(1, 5)
      PCF_ASIA_CASE(1, 6)
      PCF_ASIA_CASE(2, 2)
      PCF_ASIA_CASE(2, 3)
      PCF_ASIA_CASE(2, 4)
      PCF_ASIA_CASE(3, 2)
      PCF_ASIA_CASE(3, 1)
      PCF_ASIA_CASE(4, 1)
      PCF_ASIA_CASE(4, 2)
      PCF_ASIA_CASE(8, 1)
#endif
      PCF_ASIA_CASE(6, 1)
      default:
        set_last_error("unknown PCF_ASIA_VARIANT");
        return PCF_EINVAL;
    }
#undef PCF_ASIA_CASE
  }
  c.launches++;
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

// ------------------------------------------------------------------------------------------------
// A rank whose host failed before it could launch its kernels tells its peers: their waits fall through at once
// (xchg.cuh) instead of running into the timeout.
__global__ void xchg_poison_kernel(PeerLink link) {
  const int t = threadIdx.x;
  if (t < link.world && t != link.rank) {
    *(volatile unsigned long long*)&link.peer[t]->poison_hi = link.call_last;
    __threadfence_system();
    *(volatile unsigned long long*)&link.peer[t]->poison_lo = link.call_first;
    __threadfence_system();
  }
}

int launch_xchg_poison(Ctx& c) {
  PeerLink l = c.link;
  l.call_first = c.call_first;
  l.call_last = c.call_last;
  xchg_poison_kernel<<<1, 32, 0, c.stream>>>(l);
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

// ------------------------------------------------------------------------------------------------
// Diagnostics: the raw generator and the normal stream, for KATs and replay dumps.
__global__ void philox_kat_kernel(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed,
                                  uint32_t* out) {
  const PhiloxKey key(seed);
  uint32_t x[4];
  philox4x32_10(key, c0, c1, c2, c3, x);
  for (int i = 0; i < 4; ++i) out[i] = x[i];
}

__global__ void normal_stream_kernel(uint64_t seed, uint32_t stream, uint64_t index0, long long count,
                                     int T, double scale, const MathTables* __restrict__ tables, double* out) {
  extern __shared__ __align__(16) unsigned char tab_smem[];
  const TableView tv = stage_tables(tables, tab_smem);
  Hoisted hc;
  hc.load();
  const PhiloxKey key(seed);
  const int blocks = (T + 1) / 2;
  const long long total = count * blocks;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    long long i = g / blocks;
    int j = (int)(g - i * blocks);
    double z0, z1;
    normal_pair(key, index0 + (uint64_t)i, (uint32_t)j, stream, tv, hc, z0, z1);
    out[i * (long long)T + 2 * j] = scale * z0;
    if (2 * j + 1 < T) out[i * (long long)T + 2 * j + 1] = scale * z1;
  }
}

int run_philox_kat(Ctx& c, const unsigned int ctr[4], const unsigned int key[2], uint32_t* d_out) {
  uint64_t seed = ((uint64_t)key[1] << 32) | key[0];
  philox_kat_kernel<<<1, 1, 0, c.stream>>>(ctr[0], ctr[1], ctr[2], ctr[3], seed, d_out);
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

int run_normal_stream(Ctx& c, uint64_t seed, uint32_t stream, uint64_t index0, long long count, int T,
                      double scale, double* d_out) {
  long long total = count * ((T + 1) / 2);
  int grid = grid_for(c, total, kBlock, 8);
  normal_stream_kernel<<<grid, kBlock, kTableSmemBytes, c.stream>>>(seed, stream, index0, count, T, scale, c.d_tables, d_out);
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

}  // namespace pcf
