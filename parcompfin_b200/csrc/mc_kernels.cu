// mc_kernels.cu -- European (a1), Asian (a3) and basket (a4+a5) Monte Carlo kernels, sm_100a FP64.
// One path per thread iteration, state in registers, Philox normals generated in-register,
// compensated (sum, sum^2) reduced thread -> warp -> block -> grid. No tensor cores: nothing here
// is a dense contraction; the bound is the FP64 pipe.
#include "common.cuh"
#include "reduce.cuh"
#include "rng.cuh"
#include <cstdlib>

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
    mc_eur_kernel<true, false, 1, 2><<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, c.d_ticket, c.d_out);
  } else {
    const char* v = getenv("PCF_EUR_VARIANT");  // <pairs per thread><CTAs per SM> (tuning knob)
    const int variant = v ? atoi(v) : 81;
#define PCF_EUR_CASE(P, B)                                                                                        \
  case P * 10 + B: {                                                                                              \
    int grid = grid_for(c, (pairs.size() + P - 1) / P, kBlock, B);                                                \
    if (small)                                                                                                    \
      mc_eur_kernel<false, true, P, B><<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, c.d_ticket, c.d_out); \
    else                                                                                                          \
      mc_eur_kernel<false, false, P, B><<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, c.d_ticket, c.d_out); \
  } break;
    switch (variant) {
      PCF_EUR_CASE(1, 4)
      PCF_EUR_CASE(2, 2)
      PCF_EUR_CASE(2, 3)
      PCF_EUR_CASE(3, 2)
      PCF_EUR_CASE(3, 3)
      PCF_EUR_CASE(4, 1)
      PCF_EUR_CASE(4, 2)
      PCF_EUR_CASE(5, 1)
      PCF_EUR_CASE(6, 1)
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
      a, c.d_tables, link, c.d_partials, c.d_ticket, c.d_out);
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
    mc_asia_replay_kernel<<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, c.d_ticket, c.d_out);
  } else {
    // launch shape: PCF_ASIA_VARIANT = <paths per thread><min blocks per SM>, e.g. "14", "23" (tuning knob)
    const char* v = getenv("PCF_ASIA_VARIANT");
    const int variant = v ? atoi(v) : 61;
#define PCF_ASIA_CASE(P, B)                                   \
  case P * 10 + B:                                            \
    if (small) launch_asia<true, P, B>(c, a, paths.size(), link);   \
    else launch_asia<false, P, B>(c, a, paths.size(), link);        \
    break;
    switch (variant) {
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
      PCF_ASIA_CASE(6, 1)
      PCF_ASIA_CASE(8, 1)
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
// a4 + a5  reference include/mvn.h:78-80 (samples = L * Z) and src/mc_eur_multi.cpp:26-33.
// L (row-major, lower) sits in constant memory: every lane reads the same L[a][k] at the same
// time, which is the constant cache's broadcast case.
__constant__ double c_L[PCF_MAX_ASSETS * PCF_MAX_ASSETS];
// General kernel: per-asset parameters are folded into the constants on the host so that every FMA of the kernel has at
// most ONE constant operand (a DFMA takes one c[bank][offset]; a second one costs an LDC and its scoreboard wait):
//   c_L[a][k]  = sigma_a * A[a][k]                      -> bt[a] = sigma_a * (A z)_a
//   c_bw[a]    = w_a * S0_a * exp((r - sigma_a^2/2) T)  -> basket = sum_a c_bw[a] * exp(bt[a])     (mc_eur_multi.cpp:30)
// The reference's basket has one sigma, one S0 and weight 1/d (spec == nullptr in run_mc_basket).
__constant__ double c_bw[PCF_MAX_ASSETS];

// Keeps a loop-invariant value in a register: without this ptxas rematerialises the hoisted polynomial coefficients
// as constant loads inside the loop (341 LDC per path in the d = 16 kernel, ADU pipe 38 % busy).
__device__ __forceinline__ double pin_reg(double v) {
  asm volatile("" : "+d"(v));
  return v;
}

// exp_table() with its two two-constant FMAs fed from pinned registers
__device__ __forceinline__ double exp_table_pinned(double x, const TableView& tv, double magic, double e5) {
  const double t = fma(x, 46.16624130844683, magic);
  const double kf = t - magic;
  double r = fma(kf, -0.02166084939249829, x);
  r = fma(kf, -7.247021293269686e-19, r);
  const uint32_t n = (uint32_t)__double2loint(t);
  const double T = tv.exp_tab[(n & 31u) * tv.stride8];
  double q = fma(r, 1.0 / 720.0, e5);
  q = fma(q, r, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  const double rq = r * q;
  const double v = fma(T, rq, T);
  const int k = (int)n >> 5;
  return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
}

struct BasketArgs {
  double E, drift, sigma;  // drift = (r - sigma^2/2) T   (no sqrt(T) anywhere: SURVEY F9)
  double wS0;              // (1/d) * S0
  int cp, d;
  long long n0, n1;
  unsigned long long seed;
  const double* Z;  // replay: Z[(n-n0)*d + a]
};

// kFull: the normal transform is a full matrix (eigen-decomposition fallback of mvn.h:72-76), not a lower triangle.
// kPaths paths per thread iteration: their Philox / Box-Muller chains are independent instruction streams in one loop
// body, which is what keeps the FP64 pipe fed with one CTA per SM (same finding as mc_asia_kernel, profiles/r1_notes.md).
// kExact: a.d == D, so the per-column and per-asset tests on a.d are compile-time true and the whole path is ONE basic
// block: ptxas can then run the eight Philox / Box-Muller chains and the sixteen exponentials side by side instead of
// one (dependent) chain per block (same finding as the tree kernel's guard-free rounds, profiles/r1_notes.md).
template <int D, bool kReplay, bool kFull, int kPaths, int kMinB, bool kExact = false>
__global__ void __launch_bounds__(kBlock, kMinB) mc_basket_kernel(BasketArgs a, const MathTables* __restrict__ tables,
                                                           PeerLink link, double* partials, unsigned int* ticket,
                                                           double* out) {
  __shared__ double smem[2 * 2 * 32];
  extern __shared__ __align__(16) unsigned char tab_smem[];
  const TableView tv = stage_tables(tables, tab_smem);
  Hoisted hc;
  hc.load();
  const double xmagic = 6755399441055744.0, xe5 = 1.0 / 120.0;
  const PhiloxKey key(a.seed);
  Comp s1, s2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long nb = a.n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; nb < a.n1; nb += stride * kPaths) {
    double bt[kPaths][D];
    long long nq[kPaths];
#pragma unroll
    for (int q = 0; q < kPaths; ++q) {
      nq[q] = (nb + q * stride < a.n1) ? nb + q * stride : nb;  // clamp: a dead slot recomputes path nb, result dropped
#pragma unroll
      for (int i = 0; i < D; ++i) bt[q][i] = 0.0;
    }
    // column sweep of the triangular product: z_k is consumed as soon as it is drawn
#pragma unroll
    for (int j = 0; j < D / 2 + (D & 1); ++j) {
      if (kExact || 2 * j < a.d) {
        double z0[kPaths], z1[kPaths];
#pragma unroll
        for (int q = 0; q < kPaths; ++q) {
          if (kReplay) {
            const double* z = a.Z + (nq[q] - a.n0) * (long long)a.d;
            z0[q] = z[2 * j];
            z1[q] = (2 * j + 1 < a.d) ? z[2 * j + 1] : 0.0;
          } else {
            normal_pair(key, (uint64_t)nq[q], (uint32_t)j, PCF_STREAM_BASKET, tv, hc, z0[q], z1[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < kPaths; ++q) {
#pragma unroll
          for (int i = kFull ? 0 : 2 * j; i < D; ++i) bt[q][i] = fma(c_L[i * PCF_MAX_ASSETS + 2 * j], z0[q], bt[q][i]);
          if (2 * j + 1 < D) {
#pragma unroll
            for (int i = kFull ? 0 : 2 * j + 1; i < D; ++i)
              bt[q][i] = fma(c_L[i * PCF_MAX_ASSETS + 2 * j + 1], z1[q], bt[q][i]);
          }
        }
      }
    }
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int q = 0; q < kPaths; ++q) {
      double basket = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i)
        if (kExact || i < a.d) basket = fma(c_bw[i], exp_table_pinned(bt[q][i], tv, xmagic, xe5), basket);  // :30
      double v = payoff(basket, a.E, a.cp);
      if (kPaths > 1 && nb + q * stride >= a.n1) v = 0.0;
      t1 += v;
      t2 = fma(v, v, t2);
    }
    s1.add(t1);
    s2.add(t2);
  }
  Comp v[2] = {s1, s2};
  grid_reduce<2>(v, smem, partials, ticket, out, &link);
}

// Equicorrelation fast path. The Cholesky factor of (1-rho) I + rho 11^T has constant columns below the
// diagonal, L[a][k] = c_k for every a > k (include/mvn.h:55-70 builds exactly this matrix), so
//   Bt[a] = (sum_{k<a} c_k z_k) + L[a][a] z_a
// and the running prefix is the SAME chain of FMAs the general row-by-row product performs (bit-identical
// result), at 2 FMAs per asset instead of (a+1). No per-path array is needed, which leaves the registers for
// kPaths independent paths per thread.
__constant__ double c_Lc[PCF_MAX_ASSETS];  // c_k  = L[k+1][k]
__constant__ double c_Ld[PCF_MAX_ASSETS];  // d_a  = L[a][a]

template <int kPaths, int kMinBlocks>
__global__ void __launch_bounds__(kBlock, kMinBlocks) mc_basket_equi_kernel(BasketArgs a, const MathTables* __restrict__ tables,
                                                                   PeerLink link, double* partials,
                                                                   unsigned int* ticket, double* out) {
  __shared__ double smem[2 * 2 * 32];
  extern __shared__ __align__(16) unsigned char tab_smem[];
  const TableView tv = stage_tables(tables, tab_smem);
  Hoisted hc;
  hc.load();
  const PhiloxKey key(a.seed);
  Comp s1, s2;
  const long long T = (long long)gridDim.x * blockDim.x;
  const double sigma = a.sigma, drift = a.drift, wS0 = a.wS0;
  const int d = a.d;
  for (long long base = a.n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; base < a.n1; base += T * kPaths) {
    double prefix[kPaths], basket[kPaths];
    uint32_t lo[kPaths], hi[kPaths];
#pragma unroll
    for (int p = 0; p < kPaths; ++p) {
      const long long n = base + p * T;
      prefix[p] = 0.0;
      basket[p] = 0.0;
      lo[p] = (uint32_t)n;
      hi[p] = (uint32_t)((uint64_t)n >> 32);
    }
    for (int j = 0; 2 * j < d; ++j) {
      const double c0 = c_Lc[2 * j], d0 = c_Ld[2 * j], c1 = c_Lc[2 * j + 1], d1 = c_Ld[2 * j + 1];
      const bool two = 2 * j + 1 < d;
#pragma unroll
      for (int p = 0; p < kPaths; ++p) {
        uint32_t x[4];
        philox4x32_10(key, lo[p], hi[p], (uint32_t)j, PCF_STREAM_BASKET, x);
        double z0, z1;
        box_muller_pair(x, tv, hc, z0, z1);
        const double b0 = fma(d0, z0, prefix[p]);
        prefix[p] = fma(c0, z0, prefix[p]);
        basket[p] = fma(wS0, exp_table(fma(sigma, b0, drift), tv), basket[p]);  // mc_eur_multi.cpp:30
        if (two) {
          const double b1 = fma(d1, z1, prefix[p]);
          prefix[p] = fma(c1, z1, prefix[p]);
          basket[p] = fma(wS0, exp_table(fma(sigma, b1, drift), tv), basket[p]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < kPaths; ++p) {
      if (base + p * T < a.n1) {
        const double v = payoff(basket[p], a.E, a.cp);
        s1.add(v);
        s2.add(v * v);
      }
    }
  }
  Comp v[2] = {s1, s2};
  grid_reduce<2>(v, smem, partials, ticket, out, &link);
}

// Grid = one wave of resident CTAs (the kernel's own occupancy, not an assumed one).
template <typename K>
static int basket_launch(Ctx& c, K kernel, const BasketArgs& a, long long paths, int per_thread, const PeerLink& link) {
  int per_sm = 0;
  PCF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, kTableSmemBytes));
  const int grid = grid_for(c, (paths + per_thread - 1) / per_thread, kBlock, per_sm > 0 ? per_sm : 1);
  kernel<<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, c.d_ticket, c.d_out);
  return PCF_OK;
}

// Launch shape of the native kernel: PCF_BASKET_GEN = <paths per thread><CTAs per SM> (tuning knob): 13 | 12.
// Measured at d = 16 (profiles/r1_notes.md, r1s_tune_basket_general.log). Guarded body, 1e9 paths: 13 -> 82 ms, 22 -> 88,
// 41 -> 92, 21 -> 102. Guard-free body (kExact), 2e8 paths: 12 -> 12.2 ms, 22 -> 12.2, 21 -> 12.4, 13 -> 12.9, 11 -> 14.4
// (guarded 13: 16.4). Default 12 for the guard-free instantiation, 13 otherwise; only these two are built (every shape
// costs 20 instantiations). The replay flavour (parity path) is built once per dimension.
template <int D>
static int launch_basket(Ctx& c, const BasketArgs& a, long long paths, bool replay, bool full, const PeerLink& link) {
  if (replay) return full ? basket_launch(c, mc_basket_kernel<D, true, true, 1, 3>, a, paths, 1, link)
                          : basket_launch(c, mc_basket_kernel<D, true, false, 1, 3>, a, paths, 1, link);
  const char* e = getenv("PCF_BASKET_GEN");
  const bool exact = a.d == D && !getenv("PCF_BASKET_GUARDED");  // A/B knob: keep the per-column tests
  const int shape = e ? atoi(e) : (exact ? 12 : 13);  // one-block body: 1 path x 2 CTAs/SM (126 registers) is fastest
#define PCF_BG(P, B)                                                                                              \
  (exact ? (full ? basket_launch(c, mc_basket_kernel<D, false, true, P, B, true>, a, paths, P, link)              \
                 : basket_launch(c, mc_basket_kernel<D, false, false, P, B, true>, a, paths, P, link))            \
         : (full ? basket_launch(c, mc_basket_kernel<D, false, true, P, B>, a, paths, P, link)                    \
                 : basket_launch(c, mc_basket_kernel<D, false, false, P, B>, a, paths, P, link)))
  switch (shape) {
    case 13: return PCF_BG(1, 3);
    case 12: return PCF_BG(1, 2);
    // (11, 21 and 22 were measured too, profiles/r1s_tune_basket_general.log; not built: 70 instantiations cost minutes)
    default:
      set_last_error("unknown PCF_BASKET_GEN");
      return PCF_EINVAL;
  }
#undef PCF_BG
}

// `spec` == nullptr: the reference's basket (one sigma, one S0, weights 1/d). Otherwise per-asset arrays of length d and
// `full` says whether L_host is a full matrix (eigen fallback) or a lower triangle.
int run_mc_basket(Ctx& c, const pcf_params& p, const double* L_host /* d*d row-major */,
                  Shard paths, const double* d_replay, const PeerLink& link, const BasketHost* spec) {
  const int d = p.assets;
  const bool full = spec && spec->full;
  double Lfold[PCF_MAX_ASSETS * PCF_MAX_ASSETS] = {0};  // sigma_a folded into row a (general kernel)
  {
    double w[PCF_MAX_ASSETS] = {0};
    for (int i = 0; i < d; ++i) {
      const double s_i = spec ? spec->sigma[i] : p.sigma;
      for (int k = 0; k < (full ? d : i + 1); ++k) {
        Lfold[i * PCF_MAX_ASSETS + k] = s_i * L_host[i * d + k];
      }
      const long double drift = ((long double)p.r - (long double)s_i * s_i / 2) * (long double)p.T;
      const long double ws0 = spec ? (long double)spec->weight[i] * spec->S0[i] : (long double)p.S0 / d;
      w[i] = (double)(ws0 * expl(drift));
    }
    PCF_CUDA(cudaMemcpyToSymbolAsync(c_bw, w, sizeof(w), 0, cudaMemcpyHostToDevice, c.stream));
  }
  PCF_CUDA(cudaMemcpyToSymbolAsync(c_L, Lfold, sizeof(Lfold), 0, cudaMemcpyHostToDevice, c.stream));
  BasketArgs a;
  a.E = p.E; a.sigma = p.sigma; a.cp = p.cp; a.d = d;
  a.drift = (p.r - p.sigma * p.sigma / 2) * p.T;
  a.wS0 = (1.0 / (double)d) * p.S0;
  a.n0 = paths.begin; a.n1 = paths.end;
  a.seed = p.seed; a.Z = d_replay;
  const bool rp = d_replay != nullptr;
  // constant columns below the diagonal (bitwise)? -> equicorrelation fast path
  bool equi = !rp && !spec && !getenv("PCF_BASKET_GENERAL");
  for (int k = 0; k < d && equi; ++k)
    for (int i = k + 2; i < d; ++i)
      if (L_host[i * d + k] != L_host[(k + 1) * d + k]) { equi = false; break; }
  if (equi) {
    double Lc[PCF_MAX_ASSETS] = {0}, Ld[PCF_MAX_ASSETS] = {0};
    for (int k = 0; k < d; ++k) {
      Ld[k] = L_host[k * d + k];
      Lc[k] = (k + 1 < d) ? L_host[(k + 1) * d + k] : 0.0;
    }
    PCF_CUDA(cudaMemcpyToSymbolAsync(c_Lc, Lc, sizeof(Lc), 0, cudaMemcpyHostToDevice, c.stream));
    PCF_CUDA(cudaMemcpyToSymbolAsync(c_Ld, Ld, sizeof(Ld), 0, cudaMemcpyHostToDevice, c.stream));
    const char* v = getenv("PCF_BASKET_VARIANT");  // <paths per thread><CTAs per SM> (tuning knob)
    const int variant = v ? atoi(v) : 61;
#define PCF_BASKET_CASE(P, B)                                                                              \
  case P * 10 + B: {                                                                                       \
    int grid = grid_for(c, (paths.size() + P - 1) / P, kBlock, B);                                         \
    mc_basket_equi_kernel<P, B><<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, \
                                                                            c.d_ticket, c.d_out);          \
  } break;
    switch (variant) {
      PCF_BASKET_CASE(1, 4)
      PCF_BASKET_CASE(2, 2)
      PCF_BASKET_CASE(2, 3)
      PCF_BASKET_CASE(3, 2)
      PCF_BASKET_CASE(4, 1)
      PCF_BASKET_CASE(4, 2)
      PCF_BASKET_CASE(6, 1)
      PCF_BASKET_CASE(8, 1)
      default:
        set_last_error("unknown PCF_BASKET_VARIANT");
        return PCF_EINVAL;
    }
#undef PCF_BASKET_CASE
    c.launches++;
    PCF_CUDA(cudaGetLastError());
    return PCF_OK;
  }
  const long long np = paths.size();
  if (d <= 2) PCF_TRY(launch_basket<2>(c, a, np, rp, full, link));
  else if (d <= 4) PCF_TRY(launch_basket<4>(c, a, np, rp, full, link));
  else if (d <= 8) PCF_TRY(launch_basket<8>(c, a, np, rp, full, link));
  else if (d <= 16) PCF_TRY(launch_basket<16>(c, a, np, rp, full, link));
  else PCF_TRY(launch_basket<32>(c, a, np, rp, full, link));
  c.launches++;
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

// ------------------------------------------------------------------------------------------------
// One-warp finisher of a peer-memory exchange: waits for every rank's flag, adds in rank order.
__global__ void xchg_finish_kernel(PeerLink link, int k, double* out) {
  __shared__ double s[kXchgVals];
  peer_gather<kXchgVals>(link, s);
  if ((int)threadIdx.x < k) out[threadIdx.x] = s[threadIdx.x];
}

int launch_xchg_finish(Ctx& c, const PeerLink& l, int k, double* d_out) {
  xchg_finish_kernel<<<1, 32, 0, c.stream>>>(l, k, d_out);
  c.launches++;
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
