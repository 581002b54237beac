// amer_kernels.cu -- the reference's own early-exercise Monte Carlo scheme (a6 + a7 + a8).
//   reference include/common.h:168-208 (pathsfinder), src/mc_amer.cpp:23-113 (backward sweep),
//   include/common.h:98-141 (3x3 inverse + mat-vec).
//
// HBM layout (per GPU, n = local path index, Nl local paths = 2 * local antithetic pairs):
//   paths  double[M][Np]   row m-1 holds S at date m (row 0 of the reference, S0, is never read by the
//                          sweep and is not stored). Pair p owns columns p and p + Nl/2, exactly the
//                          reference's antithetic halves; both stores of a warp are 256 B contiguous.
//   when   WT[Np]          exercise date (exercise_when, mc_amer.cpp:23); the top bit is set when the cash flow
//                          was booked by the regression branch (mc_amer.cpp:100-103), which books
//                          payoff(S - E, E) [SURVEY F1]. WT = uint8 for M <= 127, uint16 otherwise: 1-2 B/path.
// That is the whole per-path state. The reference's discounted cash flow is exp(-r dt (when-m)) *
// payoff(paths[when][n]) (mc_amer.cpp:50), a row gather per in-the-money path and date; exercise_st is a pure
// function of that payoff and the flag (st = flag ? payoff(cp*cash, E) : cash, because cp*cash == S - E exactly
// for an in-the-money path). The sweep kernel of date m has row m in registers, so a path that exercised AT m
// -- the common case -- needs no gather; only paths whose exercise date is older re-read paths[when][n]
// (one 32-byte sector, shared by neighbours with the same date). Round-1 builds carried the gathered value in a
// `cash[n]` array instead: 16 B/path-date of extra read+write traffic, 38.8 B/path-date in total against ~18 B now.
//
// Rows and `when` are padded to a whole number of sweep tiles (Np, a multiple of 2944 paths in the shipped shape) with
// never-in-the-money dummies, so that every TMA bulk copy is a complete tile and no quad needs a bounds test.
//
// Backward sweep: one kernel per exercise date m = M .. 1, each launched as a programmatic dependent of the one before.
// The kernel of date m (a) takes the moments of date m (the previous launch's sums on one GPU, every rank's
// publication in this GPU's NVLink mailbox on several, the all-reduced buffer on the NCCL fallback path) and solves the
// 3x3 normal equations in the reference's operation order without FMA contraction, (b) applies the exercise decision of
// date m to the dates of each tile, (c) accumulates the regression moments of date m-1 from that updated state and row
// m-1, and (d) its last CTA folds the per-CTA sums in block order and publishes them. Date 1 accumulates the final
// discounted sum (mc_amer.cpp:109-111) instead of (c); date M decides nothing (mc_amer.cpp:23-27).
#include "common.cuh"
#include "reduce.cuh"
#include "rng.cuh"
#include <algorithm>
#include <type_traits>
#include <cstdlib>

namespace pcf {

constexpr int kAmerBlock = 256;
constexpr int kMaxDates = 2048;  // discount tables: constant memory -> staged into shared memory per block

// Per-call tables, ONE upload: [0, 64) the path kernel's (e^a 2^(j/32), e^a 2^(-j/32)) pairs, then M+1 factors
// exp(-r*dt*k) as mc_amer.cpp:50 evaluates them, then M+1 factors exp(-r*k*dt) as mc_amer.cpp:110 does.
constexpr int kDiscFwd = 2 * kExpEntries;
__constant__ double c_amer_tab[kDiscFwd + 2 * (kMaxDates + 1)];

struct AmerArgs {
  double S0, E;
  double adt;   // (r - sigma^2/2) dt
  double cs;    // sigma*sd (native) | sigma (replay)
  int cp, M;
  long long p0;       // first global pair of this GPU
  long long H;        // local pairs; Nl = 2H
  long long Np;       // padded row length (multiple of 4)
  unsigned long long seed;
  const double* w;    // replay: w[(p-p0)*M + (m-1)]
};

// a6: antithetic pairs, S+ and S- in registers, one Philox block per two dates, kPairs pairs per thread.
// exp((r-s^2/2)dt +- s w): x = s w is reduced once, x = (32k + j) ln2/32 + r, and
//   e^{a+x} = 2^k  [e^a 2^{ j/32}] (C(r) + S(r)),   e^{a-x} = 2^-k [e^a 2^{-j/32}] (C(r) - S(r))
// with C/S the even/odd parts of e^r (degree 6/5; |r| <= ln2/64). The bracketed factors come from a per-call
// 32-entry table (c_amer_tab[0..63], e^a folded in on the host), one LDS.128 per step: 15 FP64 for both exponentials.

__device__ __forceinline__ void amer_step(double& Sp, double& Sm, double z, double cs, const Pair* __restrict__ s_T) {
  const double x = cs * z;
  const double magic = 6755399441055744.0;
  const double t = fma(x, 46.16624130844683, magic);
  const double kf = t - magic;
  double r = fma(kf, -0.02166084939249829, x);
  r = fma(kf, -7.247021293269686e-19, r);
  const uint32_t n = (uint32_t)__double2loint(t);
  const Pair T = s_T[(n & 31u) * kRep16];
  const double r2 = r * r;
  double ce = fma(r2, 1.0 / 720.0, 1.0 / 24.0);
  ce = fma(ce, r2, 0.5);
  ce = fma(ce, r2, 1.0);
  double so = fma(r2, 1.0 / 120.0, 1.0 / 6.0);
  so = fma(so, r2, 1.0);
  so *= r;
  const double ep = T.x * (ce + so), em = T.y * (ce - so);
  const int k = (int)n >> 5;
  // scale by 2^(+-k) through the exponent field (|k| is small: |x| = sigma sqrt(dt) |z|)
  const double fp = __hiloint2double(__double2hiint(ep) + (k << 20), __double2loint(ep));
  const double fm = __hiloint2double(__double2hiint(em) - (k << 20), __double2loint(em));
  Sp *= fp;  // common.h:202
  Sm *= fm;  // common.h:203
}

template <bool kReplay, int kPairs, int kMinBlocks>
__global__ void __launch_bounds__(kAmerBlock, kMinBlocks) amer_paths_kernel(AmerArgs a, const MathTables* __restrict__ tables,
                                                                   double* __restrict__ paths) {
  extern __shared__ __align__(16) unsigned char tab_smem[];
  const TableView tv = stage_tables(tables, tab_smem);
  Pair* s_T = reinterpret_cast<Pair*>(tab_smem + kTableSmemBytes);
  for (int i = threadIdx.x; i < kExpEntries * kRep16; i += blockDim.x) {
    s_T[i].x = c_amer_tab[2 * (i / kRep16)];
    s_T[i].y = c_amer_tab[2 * (i / kRep16) + 1];
  }
  __syncthreads();
  const Pair* my_T = s_T + (threadIdx.x & (kRep16 - 1));
  Hoisted hc;
  hc.load();
  const PhiloxKey key(a.seed);
  const long long Np = a.Np;  // row stride
  const long long T = (long long)gridDim.x * blockDim.x;
  const double cs = a.cs;
  for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < a.H; base += T * kPairs) {
    double Sp[kPairs], Sm[kPairs];
    long long pp[kPairs];
#pragma unroll
    for (int q = 0; q < kPairs; ++q) {
      Sp[q] = a.S0;
      Sm[q] = a.S0;
      pp[q] = (base + q * T < a.H) ? base + q * T : base;  // clamp: duplicates rewrite identical values
    }
    if (kReplay) {
      for (int m = 1; m <= a.M; ++m) {
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
          amer_step(Sp[q], Sm[q], a.w[pp[q] * (long long)a.M + (m - 1)], cs, my_T);
          __stcs(paths + (size_t)(m - 1) * Np + pp[q], Sp[q]);
          __stcs(paths + (size_t)(m - 1) * Np + pp[q] + a.H, Sm[q]);
        }
      }
    } else {
      // Two dates per Philox block. The loop body carries no test on m, so it is ONE basic block in which ptxas runs
      // the kPairs Philox / Box-Muller chains side by side (with a per-date test every pair was its own block and its
      // dependent chain ran alone: 10.2 ms instead of the 7.4 ms the FP64 pipe needs at 1e8 x 50); an odd M ends
      // with one single-date step.
      int m = 1;
      for (; m + 1 <= a.M; m += 2) {
        double z0[kPairs], z1[kPairs];
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
          const uint64_t g = (uint64_t)(a.p0 + pp[q]);
          uint32_t x[4];
          philox4x32_10(key, (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)((m - 1) >> 1), PCF_STREAM_AMER, x);
          box_muller_pair(x, tv, hc, z0[q], z1[q]);
        }
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
          amer_step(Sp[q], Sm[q], z0[q], cs, my_T);
          __stcs(paths + (size_t)(m - 1) * Np + pp[q], Sp[q]);
          __stcs(paths + (size_t)(m - 1) * Np + pp[q] + a.H, Sm[q]);
        }
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
          amer_step(Sp[q], Sm[q], z1[q], cs, my_T);
          __stcs(paths + (size_t)m * Np + pp[q], Sp[q]);
          __stcs(paths + (size_t)m * Np + pp[q] + a.H, Sm[q]);
        }
      }
      if (m <= a.M) {
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
          const uint64_t g = (uint64_t)(a.p0 + pp[q]);
          uint32_t x[4];
          philox4x32_10(key, (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)((m - 1) >> 1), PCF_STREAM_AMER, x);
          double z0, z1;
          box_muller_pair(x, tv, hc, z0, z1);
          amer_step(Sp[q], Sm[q], z0, cs, my_T);
          __stcs(paths + (size_t)(m - 1) * Np + pp[q], Sp[q]);
          __stcs(paths + (size_t)(m - 1) * Np + pp[q] + a.H, Sm[q]);
        }
      }
    }
  }
}

// Padding columns [2H, Np): S chosen so that payoff == 0 at every date (never in the money, never gathered).
__global__ void amer_pad_kernel(double* __restrict__ paths, long long Nl, long long Np, int M, int cp) {
  const long long pad = Np - Nl;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pad * M;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / pad, col = Nl + i % pad;
    paths[row * Np + col] = cp > 0 ? 0.0 : 1e300;  // call: S - E < 0;  put: E - S < 0
  }
}

// a8: include/common.h:98-141 in the reference's operation order (cyclic %3 indexing, adjugate /
// determinant, then row-by-row mat-vec accumulated from 0). Returns false when det <= 0.
__device__ bool solve3_reference_order(const double* mom, double coef[3]) {
  const double x[3][3] = {{mom[0], mom[1], mom[2]}, {mom[1], mom[2], mom[3]}, {mom[2], mom[3], mom[4]}};
  const double y[3] = {mom[5], mom[6], mom[7]};
  double det = 0.0;
  for (int i = 0; i < 3; ++i) {
    double t = __dadd_rn(__dmul_rn(x[1][(i + 1) % 3], x[2][(i + 2) % 3]),
                         -__dmul_rn(x[1][(i + 2) % 3], x[2][(i + 1) % 3]));
    det = __dadd_rn(det, __dmul_rn(x[0][i], t));
  }
  if (!(det > 0.0)) return false;
  double inv[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double t = __dadd_rn(__dmul_rn(x[(j + 1) % 3][(i + 1) % 3], x[(j + 2) % 3][(i + 2) % 3]),
                           -__dmul_rn(x[(j + 1) % 3][(i + 2) % 3], x[(j + 2) % 3][(i + 1) % 3]));
      inv[j][i] = __ddiv_rn(t, det);
    }
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
    for (int j = 0; j < 3; ++j) s = __dadd_rn(s, __dmul_rn(inv[i][j], y[j]));
    coef[i] = s;
  }
  return true;
}

// ---- a7 (mc_amer.cpp:41-106): the backward sweep --------------------------------------------------------------
// amer_sweep_kernel, one launch per date. A CTA is 23 consumer warps + 1 producer warp, ONE CTA per SM (24 warps let
// ptxas have 80 registers, which the per-quad body needs to stay out of local memory; with three 8+1-warp CTAs per SM
// at 72 registers the three drifted apart and every date ended with SMs running one CTA: 38.9 -> 33.6 ms at 1e8 x 50,
// profiles/r2_tune_amer_chain_shapes.log). The producer's elected lane streams tiles of row m, row m-1 and the dates
// through a 3-deep mbarrier ring with cp.async.bulk; tile k of CTA b is b + k*grid, so the sums are bit-reproducible.
// Also measured, and not kept: a persistent all-dates kernel with an in-kernel grid barrier (slower on every size,
// profiles/r2_tune_amer_persistent_deferred_gathers.log); tiles handed out on demand from an atomic counter (as fast as
// the one-CTA shape, but the assignment -- hence the rounding of the sums -- changes from run to run; kept behind
// PCF_AMER_ONDEMAND in tuning builds).
// mom[0..7] = n_itm, Sx, Sx^2, Sx^3, Sx^4, Sy, Syx, Syx^2 with x = S - E, y = discounted cash flow; products are formed
// like the reference forms them unless kContractMoments folds the last multiplication into the sum's FMA.
constexpr int kMomFold = 16;  // tiles (64 paths per thread) between folds of the plain running sums
#ifdef PCF_EXACT_MOMENT_PRODUCTS
constexpr bool kContractMoments = false;  // every moment product rounded as the reference rounds it (mc_amer.cpp:51-57)
#else
constexpr bool kContractMoments = true;
#endif
// Launch shape of the sweep: kWarps consumer warps (4 paths of every tile per consumer thread) + the producer warp,
// kCtas CTAs per SM
template <int kWarps_, int kCtas_>
struct SweepShape {
  static constexpr int kWarps = kWarps_, kCtas = kCtas_;
  static constexpr int kConsumers = 32 * kWarps;
  static constexpr int kBlock = kConsumers + 32;
  static constexpr int kTile = 4 * kConsumers;  // paths per tile
};
constexpr int kMaxStages = 6;

#ifdef PCF_TUNING
// timing-only experiments (results are WRONG with any bit set): 1 no gathers, 2 no per-path work at all (the consumers
// only pull the tile out of the ring), 4 no stores of exercise dates, 8 deferred gathers are issued but not waited for (tools/tune_amer_chain.py)
__constant__ int c_sweep_knobs;
#endif

template <typename WT> struct WhenBits;
template <> struct WhenBits<uint8_t> { static constexpr int kFlag = 0x80, kMask = 0x7f; };
template <> struct WhenBits<uint16_t> { static constexpr int kFlag = 0x8000, kMask = 0x7fff; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded by the nanosecond timer (3x the exchange timeout: a producer legitimately waits while its consumers sit in a
// date barrier): a protocol error must end in a trapped kernel -- an error code on the host -- never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 0) {
  uint32_t done;
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  do {
    if (hint_ns) {
      // suspend-time hint: the thread sleeps in hardware until the phase completes (or hint_ns pass) instead of spinning
      // through issue slots (used by the producer lane, whose try_wait loop was 12 % of all executed instructions)
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
          : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity)
          : "memory");
    }
    if (!done && (++spins & 63u) == 0u) {
      const unsigned long long now = xchg_now_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 3ull * kXchgTimeoutNs) __trap();
    }
  } while (!done);
}
// Consumer side of the ring, on 32-bit shared-state-space addresses computed once per kernel (a generic pointer makes the
// compiler rebuild the shared-window address -- S2UR / UMOV / LEA -- at every use inside the tile loop). The first
// probe is inline; only a phase that is not yet complete enters the bounded loop.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  uint32_t done, spins = 0;
  unsigned long long t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (++spins & 63u) == 0u) {
      const unsigned long long now = xchg_now_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 3ull * kXchgTimeoutNs) __trap();
    }
  } while (!done);
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  if (!done) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ long long lds_s64(uint32_t a) {
  long long v;
  asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t opaque(uint32_t x) {  // keeps a loop-invariant address in its register
  asm volatile("" : "+r"(x));
  return x;
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`; streamed data is marked evict-first in L2
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
template <int kConsumers>
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }

template <typename WT>
__global__ void amer_fill_when_kernel(WT* __restrict__ when, long long Np, int M) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < Np; i += (long long)gridDim.x * blockDim.x)
    when[i] = (WT)M;
}

// Exercise dates of one quad of paths: uint8 x 4 (one 32-bit word) or uint16 x 4 (one 64-bit word).
template <typename WT> struct WhenQuad;
template <> struct WhenQuad<uint8_t> {
  typedef uint32_t Vec;
  static __device__ __forceinline__ void unpack(Vec v, int (&w)[4]) {
    w[0] = v & 0xff; w[1] = (v >> 8) & 0xff; w[2] = (v >> 16) & 0xff; w[3] = v >> 24;
  }
  static __device__ __forceinline__ Vec pack(const int (&w)[4]) {
    return (uint32_t)w[0] | ((uint32_t)w[1] << 8) | ((uint32_t)w[2] << 16) | ((uint32_t)w[3] << 24);
  }
  static __device__ __forceinline__ Vec splat(int d) { return 0x01010101u * (uint32_t)d; }
  static __device__ __forceinline__ bool same(Vec a, Vec b) { return a == b; }
  static __device__ __forceinline__ Vec lds(uint32_t a) {
    Vec v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
  }
};
template <> struct WhenQuad<uint16_t> {
  typedef uint2 Vec;
  static __device__ __forceinline__ void unpack(Vec v, int (&w)[4]) {
    w[0] = v.x & 0xffff; w[1] = v.x >> 16; w[2] = v.y & 0xffff; w[3] = v.y >> 16;
  }
  static __device__ __forceinline__ Vec pack(const int (&w)[4]) {
    return make_uint2((uint32_t)w[0] | ((uint32_t)w[1] << 16), (uint32_t)w[2] | ((uint32_t)w[3] << 16));
  }
  static __device__ __forceinline__ Vec splat(int d) { return make_uint2(0x00010001u * (uint32_t)d, 0x00010001u * (uint32_t)d); }
  static __device__ __forceinline__ bool same(Vec a, Vec b) { return a.x == b.x && a.y == b.y; }
  static __device__ __forceinline__ Vec lds(uint32_t a) {
    Vec v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
  }
};

// What every consumer thread needs to know about date m once its regression is solved.
//   mode 0 skip (no path in the money, mc_amer.cpp:73), 1 few-paths branch (:75-83), 2 regression branch with the
//   reference's rule (:97-106), 3 regression branch with the textbook rule (PCF_FLAG_AMER_LSM), 4 the reference's rule
//   on a date where it provably exercises EVERY in-the-money path (make_rule)
struct DateRule {
  int mode, booked;
  double c0, c1s, c2, nE2, sentinel;
};

// Thread 0 of a CTA: moments of date m -> s_mode / s_coef (mc_amer.cpp:73-95, common.h:98-141)
__device__ __forceinline__ void solve_date(const double* s_mom, int lsm, int* s_mode, double* s_coef, int* err_flag) {
  const double cnt = s_mom[0];
  if (cnt == 0.0) {
    *s_mode = 0;                       // mc_amer.cpp:73
  } else if (cnt <= 2.0) {
    *s_mode = 1;                       // mc_amer.cpp:75
  } else {
    double coef[3];
    if (solve3_reference_order(s_mom, coef)) {
      *s_mode = lsm ? 3 : 2;
      s_coef[0] = coef[0]; s_coef[1] = coef[1]; s_coef[2] = coef[2];
    } else {
      *s_mode = 0;
      if (blockIdx.x == 0) *(volatile int*)err_flag = PCF_ESINGULAR;  // common.h:115-117 (host-mapped status word)
    }
  }
}

template <typename WT>
__device__ __forceinline__ DateRule make_rule(int mode, const double* s_coef, double sgn, double nE, int m) {
  // Everything works on cx = cp*(S - E) = fma(sgn, S, -sgn*E): bit-identical to the reference's (double)cp*(S - E)
  // (negation is exact and rounding is symmetric), one DFMA. The regressor x = S - E is sgn*cx, so x^2, x^4 and y x^2
  // are sign-free and Sx, Sx^3, Syx are sgn times the sums formed from cx.
  // reference rule (mode 2): exercise when payoff(x, E) = max(cp*(x - E), 0) = max(cx + nE, 0) exceeds the fit
  // (mc_amer.cpp:100), and a path with x == -1 collides with the reference's sentinel and is skipped
  // (mc_amer.cpp:32,98). PCF_FLAG_AMER_LSM (mode 3): the true payoff cx, no sentinel, books without the flag.
  DateRule R;
  R.mode = mode;
  R.c0 = s_coef[0];
  R.c1s = s_coef[1] * sgn;  // c1*x == (c1*sgn)*cx exactly
  R.c2 = s_coef[2];
  R.nE2 = (mode == 2) ? nE : 0.0;
  R.sentinel = (mode == 2) ? -sgn : __longlong_as_double(0x7ff8000000000000LL);
  R.booked = (mode == 2) ? (m | WhenBits<WT>::kFlag) : m;
  // The reference's rule on a put compares payoff(x, E) = E + cx with the fit. An in-the-money path has
  // cx = E - S in (0, E] (S > 0), so E + cx > E, while |fit| <= |c0| + |c1| E + |c2| E^2 =: B on that interval whatever
  // the rounding (a relative 1e-9 covers the five roundings of the fit a million times over). Whenever B < E -- every
  // date of BASELINE config 5, where the fit is the continuation value of an at-the-money put, an order of magnitude
  // below E -- the test is true for every in-the-money path and the polynomial need not be evaluated per path: the
  // decision is exactly the reference's, with 7 FP64 instructions per path fewer. Calls and the LSM rule never qualify.
  if (mode == 2 && sgn < 0.0 && nE > 0.0) {
    const double B = fabs(R.c0) + fabs(R.c1s) * nE + fabs(R.c2) * nE * nE;
    if (B * (1.0 + 1e-9) < nE) R.mode = 4;
  }
  return R;
}


// exercise test of one path (mc_amer.cpp:97-103): returns `booked` when the path exercises at this date, else `w`
__device__ __forceinline__ int decide_regression(double cx, const DateRule& R, int w) {
  const double yhat = __dadd_rn(__dadd_rn(R.c0, __dmul_rn(R.c1s, cx)), __dmul_rn(R.c2, __dmul_rn(cx, cx)));
  const double pq = __dadd_rn(cx, R.nE2);  // before the max(., 0): max(t, 0) > y <=> t > y | 0 > y
  int out;
  asm("{\n\t.reg .pred p, q;\n\t"
      "setp.gt.f64 p, %1, 0d0000000000000000;\n\t"
      "setp.neu.and.f64 p, %1, %2, p;\n\t"
      "setp.lt.f64 q, %3, 0d0000000000000000;\n\t"
      "setp.gt.or.f64 q, %4, %3, q;\n\t"
      "and.pred p, p, q;\n\t"
      "selp.s32 %0, %5, %6, p;\n\t}"
      : "=r"(out)
      : "d"(cx), "d"(R.sentinel), "d"(yhat), "d"(pq), "r"(R.booked), "r"(w));
  return out;
}

// One quad (4 consecutive paths, one 32-byte sector of row 1) at the LAST step of the sweep, date m = 1: the decision of
// date 1 on the dates w[], then the discounted cash flows (mc_amer.cpp:109-111) added to run[0] (sum) and run[1] (sum
// of squares). Returns true when any of the four dates changed. colp + d*row_bytes is the address of paths[d][first
// path of the quad]; s_disc[k] = exp(-r dt k), s_abs[k] = exp(-r k dt).
template <typename WT>
__device__ __forceinline__ bool sweep_quad_final(const double (&src)[4], int (&w)[4], const DateRule& R, int m,
                                                 double sgn, double nE, const char* colp, size_t row_bytes,
                                                 const double* s_disc, const double* s_abs, double (&run)[8]) {
  constexpr int kFlag = WhenBits<WT>::kFlag, kMask = WhenBits<WT>::kMask;
  double cx[4], gv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    cx[e] = fma(sgn, src[e], nE);  // payoff(S_m) = max(cx, 0)
    gv[e] = src[e];
  }
  // (b) decision of date m
  bool changed = false;
  if (R.mode >= 2) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int wn = decide_regression(cx[e], R, w[e]);
      changed |= wn != w[e];
      w[e] = wn;
    }
  } else if (R.mode == 1) {
    // <= 2 paths in the money (mc_amer.cpp:75-83): true payoff against the discounted cash flow
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (cx[e] > 0.0) {
        const int d = w[e] & kMask;
        const double g = __ldg(reinterpret_cast<const double*>(colp + (size_t)(unsigned)d * row_bytes) + e);
        const double pg = fma(sgn, g, nE);  // payoff(g, E, cp) = max(cp*(g - E), 0), same rounding
        const double cont = __dmul_rn(s_disc[d - m], pg > 0.0 ? pg : 0.0);
        if (cx[e] > cont) {
          w[e] = m;
          changed = true;
        }
      }
    }
  }
  // (c) the cash flow of a path that did not exercise at m comes from paths[when][n] (mc_amer.cpp:50)
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int d = w[e] & kMask;
    if (d != m) gv[e] = __ldg(reinterpret_cast<const double*>(colp + (size_t)(unsigned)d * row_bytes) + e);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int d = w[e] & kMask;
    const double cr = fma(sgn, gv[e], nE);
    const double cs = cr > 0.0 ? cr : 0.0;  // payoff(paths[when][n])
    // exercise_st (mc_amer.cpp:103): the regression branch booked payoff(x, E) = max(cp*(x - E), 0), x = S - E;
    // a flagged path has cs > 0
    const double sq = __dadd_rn(cs, nE);
    const double stv = (w[e] & kFlag) ? (sq > 0.0 ? sq : 0.0) : cs;
    const double v = __dmul_rn(s_abs[d], stv);
    run[0] += v;
    run[1] += v * v;
  }
  return changed;
}

// ---- dates M .. 2: the lean per-quad body -------------------------------------------------------------------------
// The reference's per-path arithmetic (mc_amer.cpp:41-59, 97-103), arranged for the instruction count: the ring alone
// streams a date in 0.22 ms, the full kernel needs 0.47 ms, and its time follows the number of instructions issued per
// tile (profiles/r2_tune_amer_chain_*.log):
//   * the regressor and the cash flow are carried DOUBLED, X = c + |c| = 2 max(c, 0) and C = cr + |cr| = 2 payoff: one
//     DADD instead of a compare and two selects. Every term of a moment is then the reference's term times an exact
//     power of two (2 x, 4 x^2, 8 x^3, 16 x^4, 2 y, 4 y x, 8 y x^2), rounded exactly as the unscaled product is, and
//     the sums are scaled back exactly when the CTA totals are formed;
//   * a path out of the money at m-1 has X = 0, which zeroes its x and y x^k terms by itself; a path that exercised AT
//     m (the common case) is one date ahead of m-1, so its discount factor is the constant exp(-r dt), no table look-up;
//   * every tile is complete (rows are padded to a multiple of the tile with never-in-the-money columns), so there is
//     no per-quad bounds test;
//   * the exercise test is four FP64 compares chained through their predicate operands (inline PTX: left to itself the
//     compiler turns max(t, 0) > y into a NaN-propagating maximum, seven instructions);
//   * ONE gather per thread and tile is deferred: it is issued here as an asynchronous copy and its terms are added
//     while the NEXT tile is processed (Pending), so the DRAM round trip of paths[when][n] -- 30 % of all warp stall
//     samples in the immediate version, profiles/r2f_ncu_amer_sweep_immediate_gathers.txt -- overlaps a whole tile of
//     work; the paths that need one are collected in a bit mask. A thread's first gather of a tile keeps its regressor
//     in registers; a second one (0.5 % of the threads, but 15-18 % of the warps have such a thread in every tile, and
//     consuming it on the spot cost 7 % of all stall samples) goes through a small record in shared memory; a third
//     or fourth is consumed at once.
struct Pending {
  uint32_t g;  // this thread's landing slot in shared memory for paths[when][n] (holds a finite value at all times)
  double X;   // 2 max(cp (S_{m-1} - E), 0) of that path
  int j;      // when - (m-1); 0: nothing pending (s_dz[0] == 0.0 zeroes the terms); kSecond: see g2
  uint32_t g2;  // a thread's SECOND gather of a tile (15 % of the warps have such a thread in every tile): a 24-byte
                // record in shared memory -- the value (asynchronous copy), X and the table index -- consumed with the
                // first one; kept out of the registers because it is rare
};
constexpr int kSecond = 1 << 20;
// The deferred value travels by cp.async, not by a load into a register: the fence that hands the ring slot back
// (fence.proxy.async -> MEMBAR.ALL.CTA) waits for every outstanding LOAD of the thread, so a register gather left in
// flight across the tile boundary stalls the slot release instead of overlapping it (measured: 38.2 ms against 33.6 ms
// at 1e8 x 50, profiles/r2_tune_amer_chain_lean.log); asynchronous copies are only waited for by cp.async.wait_all.
__device__ __forceinline__ void cp_async8(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ double pending_value(const Pending& pend) {
#ifdef PCF_TUNING
  if (!(c_sweep_knobs & 8))  // timing experiment: the deferred value is used without waiting for it
#endif
  asm volatile("cp.async.wait_all;" ::: "memory");
  return lds_f64(pend.g);
}

template <bool kContract>
__device__ __forceinline__ void x_terms(double X, double X2, double (&run)[8]) {
  run[1] += X;
  run[2] += X2;
  if (kContract) {
    run[3] = fma(X2, X, run[3]);
    run[4] = fma(X2, X2, run[4]);
  } else {
    const double X3 = __dmul_rn(X2, X);
    run[3] += X3;
    run[4] += __dmul_rn(X3, X);
  }
}
template <bool kContract>
__device__ __forceinline__ void y_terms(double Y, double X, double X2, double (&run)[8]) {
  run[5] += Y;
  if (kContract) {
    run[6] = fma(Y, X, run[6]);
    run[7] = fma(Y, X2, run[7]);
  } else {
    const double YX = __dmul_rn(Y, X);
    run[6] += YX;
    run[7] += __dmul_rn(YX, X);
  }
}

// The terms of the gathers a thread deferred in the previous tile (all lanes of the warp call it)
__device__ __forceinline__ void consume_pending(Pending& pend, double sgn, double nE, uint32_t s_dz, double (&run)[8]) {
  const double cr = fma(sgn, pending_value(pend), nE);
  const double Y = __dmul_rn(lds_f64(s_dz + 8u * (uint32_t)(pend.j & (kSecond - 1))), cr + fabs(cr));
  y_terms<kContractMoments>(Y, pend.X, __dmul_rn(pend.X, pend.X), run);
  if (__any_sync(0xffffffffu, pend.j & kSecond)) {
    if (pend.j & kSecond) {
      const double cr2 = fma(sgn, lds_f64(pend.g2), nE);
      const double X2 = lds_f64(pend.g2 + 8u);
      const double Y2 = __dmul_rn(lds_f64(s_dz + 8u * lds_u32(pend.g2 + 16u)), cr2 + fabs(cr2));
      y_terms<kContractMoments>(Y2, X2, __dmul_rn(X2, X2), run);
    }
  }
  pend.j = 0;
}

// s_dz[j] = exp(-r dt j) for j >= 1, s_dz[0] = 0.0; disc1 = s_dz[1]. colp + d*row_bytes is the address of
// paths[d][first path of the quad].
template <typename WT>
__device__ __forceinline__ void sweep_quad_moments(const double (&src)[4], const double (&sp)[4], int (&w)[4],
                                                   const DateRule& R, int m, double sgn, double nE, double disc1,
                                                   const char* colp, size_t row_bytes, uint32_t s_dz,
                                                   double (&run)[8], int& cnt, Pending& pend) {
  constexpr int kMask = WhenBits<WT>::kMask;
  double cx[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) cx[e] = fma(sgn, src[e], nE);  // cp (S_m - E): payoff(S_m) = max(cx, 0)
  // (b) decision of date m
  if (R.mode == 4) {
#pragma unroll
    for (int e = 0; e < 4; ++e) w[e] = (cx[e] > 0.0 && cx[e] != R.sentinel) ? R.booked : w[e];
  } else if (R.mode >= 2) {
#pragma unroll
    for (int e = 0; e < 4; ++e) w[e] = decide_regression(cx[e], R, w[e]);
  } else if (R.mode == 1) {
    // <= 2 paths in the money (mc_amer.cpp:75-83): true payoff against the discounted cash flow
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (cx[e] > 0.0) {
        const int d = w[e] & kMask;
        const double g = __ldg(reinterpret_cast<const double*>(colp + (size_t)(unsigned)d * row_bytes) + e);
        const double pg = fma(sgn, g, nE);  // payoff(g, E, cp) = max(cp*(g - E), 0), same rounding
        const double cont = __dmul_rn(lds_f64(s_dz + 8u * (uint32_t)(d - m)), pg > 0.0 ? pg : 0.0);  // d > m: never entry 0
        if (cx[e] > cont) w[e] = m;
      }
    }
  }
  // (c) + (d) moments of date m-1. A path in the money at m-1 (mc_amer.cpp:44) whose exercise date is m takes its cash
  // flow from row m, one date ahead; any other date means paths[when][n] (mc_amer.cpp:50). The paths that need that
  // gather are collected in a bit mask, not in predicates (seven predicate registers do not hold four paths' worth of
  // conditions: the first version of this loop spent 30 instructions per tile moving predicates in and out of a GPR)
  unsigned gm = 0;
  double X[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const double c = fma(sgn, sp[e], nE);
    const bool need = c > 0.0;
    X[e] = c + fabs(c);
    const double X2 = __dmul_rn(X[e], X[e]);
    const bool at_m = (w[e] & kMask) == m;
    gm |= (need && !at_m) ? (1u << e) : 0u;
    cnt += need ? 1 : 0;
    x_terms<kContractMoments>(X[e], X2, run);
    const double Y = __dmul_rn(disc1, cx[e] + fabs(cx[e]));
    y_terms<kContractMoments>((need && at_m) ? Y : 0.0, X[e], X2, run);
  }
  // the gathers deferred by the previous tile, placed after this tile's moments (measured 1.5 % faster than at the top
  // of the tile; ptxas still schedules the wait itself -- DEPBAR has no register operand -- right after the decision, so
  // the copies get the ring wait, the slot hand-over and the decision of this tile to land, not the whole tile)
  consume_pending(pend, sgn, nE, s_dz, run);
#ifdef PCF_TUNING
  if (c_sweep_knobs & 1) return;
#endif
  // the gathers: the first one of the thread is deferred, further ones (rare) are consumed on the spot
  if (gm) {
    const int esel = __ffs(gm) - 1;
    int dsel = (esel == 0 ? w[0] : esel == 1 ? w[1] : esel == 2 ? w[2] : w[3]) & kMask;
#ifdef PCF_TUNING
    if (c_sweep_knobs & 16) dsel = m;  // timing experiment: the gather reads the row that has just been streamed
#endif
    cp_async8(pend.g, reinterpret_cast<const double*>(colp + (size_t)(unsigned)dsel * row_bytes) + esel);
    pend.X = esel == 0 ? X[0] : esel == 1 ? X[1] : esel == 2 ? X[2] : X[3];
    pend.j = dsel - (m - 1);
    unsigned rest = gm & (gm - 1);
    if (rest) {
      // second gather: deferred through the record; third and fourth (vanishingly rare) consumed on the spot
      const int e2 = __ffs(rest) - 1;  // 1..3
      const int d2 = (e2 == 1 ? w[1] : e2 == 2 ? w[2] : w[3]) & kMask;
      cp_async8(pend.g2, reinterpret_cast<const double*>(colp + (size_t)(unsigned)d2 * row_bytes) + e2);
      sts_f64(pend.g2 + 8u, e2 == 1 ? X[1] : e2 == 2 ? X[2] : X[3]);
      sts_u32(pend.g2 + 16u, (uint32_t)(d2 - (m - 1)));
      pend.j |= kSecond;
      rest &= rest - 1;
#pragma unroll
      for (int e = 2; e < 4; ++e) {
        if (rest >> e & 1u) {
          const int d = w[e] & kMask;
          const double gv = __ldg(reinterpret_cast<const double*>(colp + (size_t)(unsigned)d * row_bytes) + e);
          const double cr = fma(sgn, gv, nE);
          const double Y = __dmul_rn(lds_f64(s_dz + 8u * (uint32_t)(d - (m - 1))), cr + fabs(cr));
          y_terms<kContractMoments>(Y, X[e], __dmul_rn(X[e], X[e]), run);
        }
      }
    }
  }
}

// Fold of the plain per-thread runs (<= 4 kMomFold terms each): summed over the warp in a fixed shuffle order and added
// to the warp's compensated totals in shared memory; only lane 0's copy is used.
template <int kSums, bool kCount>
__device__ __forceinline__ void fold_runs(double (&run)[8], int& cnt, double2 (*s_wacc)[8], int tid) {
  if (kCount) run[0] = (double)cnt;
  cnt = 0;
#pragma unroll
  for (int k = 0; k < kSums; ++k) {
    double v = run[k];
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, dlt));
    if ((tid & 31) == 0) {
      const double2 t = s_wacc[tid >> 5][k];
      Comp c(t.x, t.y);
      c.add(v);
      s_wacc[tid >> 5][k] = make_double2(c.hi, c.lo);
    }
    run[k] = 0.0;
  }
}

// one ring slot: a tile of row m, of row m-1 and of the exercise dates
template <typename WT, int kTile>
__host__ __device__ constexpr size_t sweep_stage_bytes() { return (size_t)kTile * (8 + 8 + sizeof(WT)); }

// ----------------------------------------------------------------------------------------------------------------
// Per-date driver (NCCL fallback path): one launch per exercise date; rows AND dates stream through the TMA ring.
struct SweepArgs {
  const double* paths;  // row m-1 = date m, stride Np
  void* when;
  long long Np;         // multiple of 16
  double E;
  int cp, m, M;
  int rev;              // tiles are walked from the far end (alternate dates: the tail of the previous kernel's
                        // row m-1 and date tiles is still in L2 when this kernel starts there)
  int first;            // date M: no decision
  int lsm;              // PCF_FLAG_AMER_LSM
  int stages;
  const double* mom_in;
  double* partials;
  unsigned int* ticket;
  double* out;          // kMoments: the 8 moments of date m-1; kFinal: sum, sumsq
  int* err_flag;
  unsigned int* tile_ctr;   // this date's tile counter (zero on entry): tiles beyond a CTA's first are handed out on
                            // demand; null: tile k of CTA b is b + k * grid
  unsigned long long* dbg;  // PCF_TUNING builds: per CTA (smid, consumer cycles in the tile loop, tiles), else null
};

template <typename WT, bool kMoments, bool kFinal, class Shape>
__global__ void __launch_bounds__(Shape::kBlock, Shape::kCtas) amer_sweep_kernel(SweepArgs a, PeerLink link) {
  typedef WhenQuad<WT> WQ;
  typedef typename WQ::Vec WVec;
  constexpr int kSweepConsumers = Shape::kConsumers, kTilePaths = Shape::kTile;
  constexpr size_t kStage = sweep_stage_bytes<WT, kTilePaths>();
  constexpr int kSums = kFinal ? 2 : 8;
  __shared__ double smem[8 * 2 * 32];
  __shared__ double s_mom[kXchgVals];
  __shared__ double s_coef[3];
  __shared__ int s_mode;
  __shared__ __align__(8) uint64_t s_full[kMaxStages], s_empty[kMaxStages];
  __shared__ long long s_tile[kMaxStages];  // tile held by each ring slot, -1: no more tiles
  __shared__ double s_pend[kMoments ? kSweepConsumers : 1];      // landing slots of the deferred gathers
  __shared__ double s_pend2[kMoments ? 3 * kSweepConsumers : 1];  // records of the second deferred gather of a thread
  __shared__ double2 s_wacc[kSweepConsumers / 32][8];
  extern __shared__ __align__(128) unsigned char dyn[];
  unsigned char* ring = dyn;                                                     // stages x kStage
  double* s_disc = reinterpret_cast<double*>(dyn + (size_t)a.stages * kStage);   // exp(-r dt k), k = 0..M
  double* s_abs = s_disc + (a.M + 1);                                            // exp(-r k dt) (kFinal)
  const int tid = threadIdx.x;
  // Programmatic dependent launch: the kernel of date m-1 may be scheduled as soon as every CTA of this one has started
  // (its CTAs become resident as ours exit, stage their tables, pull the rows of their first tiles and park in
  // griddepcontrol.wait until this grid has completed and its stores -- dates, moments -- are visible)
  asm volatile("griddepcontrol.launch_dependents;");
  if (tid < (kSweepConsumers / 32) * 8) s_wacc[tid / 8][tid % 8] = make_double2(0.0, 0.0);
  for (int k = tid; k <= a.M; k += blockDim.x) {
    // dates M..2 index this table from 1 (a cash flow lies at least one date ahead) and read entry 0 as "no term"
    s_disc[k] = (kMoments && k == 0) ? 0.0 : c_amer_tab[kDiscFwd + k];
    s_abs[k] = c_amer_tab[kDiscFwd + (a.M + 1) + k];
  }
  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], kSweepConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int m = a.m, cp = a.cp;
  const double E = a.E;
  const long long Np = a.Np;
  const long long ntiles = Np / kTilePaths;  // Np is a multiple of the tile
  const double* __restrict__ paths = a.paths;
  const double* row_m = paths + (size_t)(m - 1) * Np;
  const double* row_p = paths + (size_t)(kMoments ? m - 2 : 0) * Np;
  WT* when = reinterpret_cast<WT*>(a.when);

  if (tid >= kSweepConsumers) {
    // ---- producer warp: one elected lane keeps the ring full
    if (tid == kSweepConsumers) {
      uint64_t pol, pol_keep;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep));
      constexpr uint32_t kPerPath = 8 + (kMoments ? 8 : 0) + sizeof(WT);
      // Tile k of CTA b is b + k*grid (shipped): a fixed assignment, hence bit-reproducible sums. With a.tile_ctr set
      // (PCF_AMER_ONDEMAND, tuning builds) the tiles after the first come off the date's atomic counter as ring slots
      // free up, requested one tile ahead so that the atomic's round trip hides behind the wait for the slot -- that
      // balances CTAs that drift apart on one SM (profiles/r2_tune_amer_chain_on_demand_pdl_envelope.log), which the
      // one-CTA-per-SM shape does not need.
      long long cur = blockIdx.x;
      auto next_tile = [&](long long prev) -> long long {
        return a.tile_ctr ? (long long)gridDim.x + (long long)atomicAdd(a.tile_ctr, 1u) : prev + (long long)gridDim.x;
      };
      auto rows = [&](int slot, long long tt) {
        const long long c0 = (a.rev ? ntiles - 1 - tt : tt) * kTilePaths;
        constexpr uint32_t n = kTilePaths;  // rows are padded to whole tiles
        unsigned char* st = ring + (size_t)slot * kStage;
        s_tile[slot] = c0;
        mbar_arrive_expect_tx(&s_full[slot], n * kPerPath);
        bulk_g2s(st, row_m + c0, n * 8u, &s_full[slot], pol);
        if (kMoments) bulk_g2s(st + kTilePaths * 8, row_p + c0, n * 8u, &s_full[slot], pol_keep);
      };
      auto dates = [&](int slot) {
        const long long c0 = s_tile[slot];
        constexpr uint32_t n = kTilePaths;
        bulk_g2s(ring + (size_t)slot * kStage + kTilePaths * 16, when + c0, n * (uint32_t)sizeof(WT), &s_full[slot], pol_keep);
      };
      // rows do not depend on the previous date's kernel: the first tiles' rows are in flight before the wait; their
      // exercise dates -- that kernel's stores -- follow it
      int npre = 0;
      while (npre < a.stages && cur < ntiles) {
        rows(npre, cur);
        cur = next_tile(cur);
        ++npre;
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int i = 0; i < npre; ++i) dates(i);
      int s = (npre == a.stages) ? 0 : npre;
      uint32_t ph = (npre == a.stages) ? 1u : 0u;
      for (;;) {
        mbar_wait(&s_empty[s], ph ^ 1);
        if (cur >= ntiles) {
          s_tile[s] = -1;
          mbar_arrive(&s_full[s]);
          break;
        }
        const long long nxt = next_tile(cur);
        rows(s, cur);
        dates(s);
        cur = nxt;
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ---- consumers
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!a.first) {
      if (link.world > 1) {
        if (tid < 32) peer_gather_warp<kXchgVals>(link, link.seq - 1ull, s_mom);  // every rank's moments of date m
      } else if (tid < kXchgVals) {
        s_mom[tid] = a.mom_in[tid];  // the previous launch's sums (all-reduced by NCCL in between when world > 1)
      }
      consumer_bar<kSweepConsumers>();
      if (tid == 0) solve_date(s_mom, a.lsm, &s_mode, s_coef, a.err_flag);
    } else if (tid == 0) {
      s_mode = 0;
    }
    consumer_bar<kSweepConsumers>();
    const double sgn = (double)cp, nE = -sgn * E;
    const DateRule R = make_rule<WT>(s_mode, s_coef, sgn, nE, m);
    double run[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) run[k] = 0.0;
    int cnt = 0, fold = 0;
    const double disc1 = c_amer_tab[kDiscFwd + 1];  // exp(-r dt): one date ahead
    Pending pend;
    pend.g = smem_u32(s_pend + (kMoments ? tid : 0));
    pend.g2 = smem_u32(s_pend2 + (kMoments ? 3 * tid : 0));
    if (kMoments) s_pend[tid] = E;
    pend.X = 0.0;
    pend.j = 0;
    const size_t row_bytes = (size_t)Np * 8;
    const char* colp0 = reinterpret_cast<const char*>(paths) + (size_t)tid * 32 - row_bytes;
    int s = 0;
    uint32_t ph = 0;
    const uint32_t full0 = opaque(smem_u32(&s_full[0])), empty0 = opaque(smem_u32(&s_empty[0]));
    const uint32_t tile0 = opaque(smem_u32(&s_tile[0])), dz0 = opaque(smem_u32(s_disc));
    pend.g = opaque(pend.g);
    pend.g2 = opaque(pend.g2);
    const uint32_t quad0 = opaque(smem_u32(ring) + (uint32_t)tid * 32u);          // this thread's sector of slot 0, row m
    const uint32_t date0 = opaque(smem_u32(ring) + (uint32_t)kTilePaths * 16u + (uint32_t)tid * (uint32_t)sizeof(WVec));
#ifdef PCF_TUNING
    const long long dbg_t0 = clock64();
    int dbg_tiles = 0;
#endif
    for (;;) {
      mbar_wait_addr(full0 + 8u * (uint32_t)s, ph);
      const long long c0t = lds_s64(tile0 + 8u * (uint32_t)s);
      if (c0t < 0) break;
      const uint32_t so = (uint32_t)s * (uint32_t)kStage;
      const double2 sa = lds_f64x2(quad0 + so);
      const double2 sb = lds_f64x2(quad0 + so + 16u);
      double2 pa = make_double2(0.0, 0.0), pb = pa;
      if (kMoments) {
        pa = lds_f64x2(quad0 + so + (uint32_t)kTilePaths * 8u);
        pb = lds_f64x2(quad0 + so + (uint32_t)kTilePaths * 8u + 16u);
      }
      const WVec wv = WQ::lds(date0 + so);
      // generic-proxy reads of the slot must be ordered before the TMA engine (async proxy) refills it: without
      // this fence a deep ring at 2 CTAs/SM produced stale reads (observed as run-to-run price noise)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive_addr(empty0 + 8u * (uint32_t)s);
      if (++s == a.stages) { s = 0; ph ^= 1; }
#ifdef PCF_TUNING
      ++dbg_tiles;
#endif

      const double src[4] = {sa.x, sa.y, sb.x, sb.y};
      const double sp[4] = {pa.x, pa.y, pb.x, pb.y};
      int w[4];
      WQ::unpack(wv, w);
      WVec* wp = reinterpret_cast<WVec*>(when + c0t) + tid;
      const char* colp = colp0 + (size_t)c0t * 8;
#ifdef PCF_TUNING
      if (c_sweep_knobs & 2) {
        run[1] += src[0] + src[1] + src[2] + src[3] + sp[0] + sp[1] + sp[2] + sp[3] + (double)w[0];
        continue;
      }
#endif
      bool changed;
      if (kMoments) {
        sweep_quad_moments<WT>(src, sp, w, R, m, sgn, nE, disc1, colp, row_bytes, dz0, run, cnt, pend);
        changed = !WQ::same(WQ::pack(w), wv);
      } else {
        changed = sweep_quad_final<WT>(src, w, R, m, sgn, nE, colp, row_bytes, s_disc, s_abs, run);
      }
#ifdef PCF_TUNING
      if (c_sweep_knobs & 4) changed = false;
#endif
      if (R.mode >= 2) {
        if (__any_sync(0xffffffffu, changed)) *wp = WQ::pack(w);  // whole 128-byte lines back
      } else if (changed) {
        *wp = WQ::pack(w);
      }
      if (++fold == kMomFold) {
        fold_runs<kSums, kMoments>(run, cnt, s_wacc, tid);
        fold = 0;
      }
    }
    if (kMoments) consume_pending(pend, sgn, nE, dz0, run);  // the last tile's deferred gathers
    fold_runs<kSums, kMoments>(run, cnt, s_wacc, tid);
#ifdef PCF_TUNING
    if (tid == 0 && a.dbg) {
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      a.dbg[blockIdx.x * 3 + 0] = smid;
      a.dbg[blockIdx.x * 3 + 1] = (unsigned long long)(clock64() - dbg_t0);
      a.dbg[blockIdx.x * 3 + 2] = (unsigned long long)dbg_tiles;
    }
#endif
  }
  // warp totals -> CTA -> grid. One warp per sum: warp k folds the per-warp totals of sum k (one lane per consumer warp,
  // fixed shuffle tree), leaves the CTA's total in global memory, and in the last CTA to arrive folds the totals of
  // all CTAs the same way (lanes stride the blocks). The generic grid_reduce (reduce.cuh) runs every sum through every
  // warp's shuffle tree twice -- 2-3 us of FP64 work per CTA and date, which a 12.5e6-path shard pays 49 times.
  __syncthreads();
  __shared__ bool s_is_last;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int kCw = kSweepConsumers / 32;
  static_assert(kCw <= 32, "one lane per consumer warp");
  for (int ks = warp; ks < kSums; ks += kCw) {
    // back from the doubled cx-space to the reference's x = S - E and y: exact powers of two, and odd powers of x carry
    // the sign of cp
    const double sgn = (double)cp;
    const double f8[8] = {1.0, 0.5 * sgn, 0.25, 0.125 * sgn, 0.0625, 0.5, 0.25 * sgn, 0.125};
    const double f = kMoments ? f8[ks] : 1.0;
    Comp v;
    if (lane < kCw) {
      const double2 t = s_wacc[lane][ks];
      v = Comp(t.x * f, t.y * f);
    }
    v = warp_reduce(v);
    if (lane == 0) {
      a.partials[((size_t)blockIdx.x * kSums + ks) * 2 + 0] = v.hi;
      a.partials[((size_t)blockIdx.x * kSums + ks) * 2 + 1] = v.lo;
      __threadfence();
    }
  }
  __syncthreads();
  if (tid == 0) s_is_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  for (int ks = warp; ks < kSums; ks += kCw) {
    constexpr int kBatch = 5;  // 5 x 32 >= 148 CTAs: one batch of independent L2 loads, then the dependent adds
    Comp acc;
    for (unsigned int b0 = lane; b0 < gridDim.x; b0 += 32 * kBatch) {
      double2 v[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const unsigned int b = b0 + 32 * j;
        v[j] = (b < gridDim.x) ? __ldcg(reinterpret_cast<const double2*>(a.partials + ((size_t)b * kSums + ks) * 2))
                               : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) acc.merge(Comp(v[j].x, v[j].y));
    }
    acc = warp_reduce(acc);
    if (lane == 0) {
      const double r = acc.value();
      a.out[ks] = r;
      smem[ks] = r;
    }
  }
  __syncthreads();
  if (tid == 0) *a.ticket = 0u;
  if (link.world > 1) {  // reduce.cuh grid_reduce: the reduction and the collective are one kernel
    peer_publish<kSums>(link, smem);
    if (link.gather) {
      __syncthreads();
      peer_gather<kSums>(link, smem + 16);
      if (tid < kSums) a.out[tid] = smem[16 + tid];
    }
  }
}

#ifdef PCF_TUNING
static unsigned long long* g_sweep_dbg = nullptr;
#endif

template <typename WT, class Shape>
static int launch_sweep(Ctx& c, bool final_date, bool dependent, int grid, int stages, const SweepArgs& a,
                        const PeerLink& link) {
  const size_t dsm = (size_t)stages * sweep_stage_bytes<WT, Shape::kTile>() + 2 * sizeof(double) * (a.M + 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(Shape::kBlock);
  cfg.dynamicSmemBytes = dsm;
  cfg.stream = c.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = dependent ? 1 : 0;
  if (final_date) {
    auto k = amer_sweep_kernel<WT, false, true, Shape>;
    PCF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
    PCF_CUDA(cudaLaunchKernelEx(&cfg, k, a, link));
  } else {
    auto k = amer_sweep_kernel<WT, true, false, Shape>;
    PCF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
    PCF_CUDA(cudaLaunchKernelEx(&cfg, k, a, link));
  }
  return PCF_OK;
}

static inline size_t amer_when_bytes(int M) { return M <= WhenBits<uint8_t>::kMask ? 1 : 2; }
static inline size_t amer_when_area(long long Np, int M) { return ((size_t)Np * amer_when_bytes(M) + 255) & ~(size_t)255; }

constexpr int kDefaultSweepShape = 231, kDefaultSweepStages = 3;
constexpr int kSweepMaxTile = 28 * 128;  // largest tile of any launch shape (tuning builds)

// Per-date chain: one launch per exercise date, each a programmatic dependent of the one before. The moments travel
// through the NVLink mailboxes (published by the last block of date m+1's kernel, gathered by every CTA of date m's),
// through `mom` on one GPU, or through an ncclAllReduce between two launches on the fallback path.
template <class Shape>
static int run_sweep_chain(Ctx& c, const pcf_params& p, const double* paths, void* when, long long Np, int stages) {
  const int M = p.M;
  const long long ntiles = (Np + Shape::kTile - 1) / Shape::kTile;
  const bool w8 = amer_when_bytes(M) == 1;
  const bool nccl_path = c.world > 1 && !use_peer(c);
  const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)c.sm_count * Shape::kCtas));
  unsigned int* tile_ctr = reinterpret_cast<unsigned int*>((char*)when + amer_when_area(Np, M));
  const bool dynamic_tiles = tuning_env("PCF_AMER_ONDEMAND") != nullptr;  // tuning builds only (not reproducible bit for bit)
  if (dynamic_tiles) PCF_CUDA(cudaMemsetAsync(tile_ctr, 0, sizeof(unsigned int) * (size_t)(M + 1), c.stream));
  {
    const int fg = grid_for(c, Np, 256, 8);
    if (w8) amer_fill_when_kernel<uint8_t><<<fg, 256, 0, c.stream>>>((uint8_t*)when, Np, M);
    else amer_fill_when_kernel<uint16_t><<<fg, 256, 0, c.stream>>>((uint16_t*)when, Np, M);
    c.launches++;
  }
  SweepArgs sa;
  sa.paths = paths; sa.when = when; sa.Np = Np; sa.E = p.E; sa.cp = p.cp; sa.M = M;
  sa.lsm = (p.flags & PCF_FLAG_AMER_LSM) ? 1 : 0;
  sa.stages = stages;
  sa.partials = c.d_partials; sa.ticket = c.d_ticket; sa.err_flag = c.flag_dev;
  sa.dbg = nullptr;
  const bool pdl = !nccl_path && tuning_env("PCF_AMER_NOPDL") == nullptr;
#ifdef PCF_TUNING
  {
    const int knobs = tuning_env("PCF_AMER_ENVELOPE") ? atoi(tuning_env("PCF_AMER_ENVELOPE")) : 0;
    PCF_CUDA(cudaMemcpyToSymbolAsync(c_sweep_knobs, &knobs, sizeof(int), 0, cudaMemcpyHostToDevice, c.stream));
  }
  if (!g_sweep_dbg) PCF_CUDA(cudaMalloc(&g_sweep_dbg, sizeof(unsigned long long) * 3 * 1024));
  const int dbg_date = tuning_env("PCF_AMER_DBG_DATE") ? atoi(tuning_env("PCF_AMER_DBG_DATE")) : M / 2;
#endif
  double* mom[2] = {c.d_out + 8, c.d_out + 16};
  for (int m = M; m >= 1; --m) {
    sa.m = m;
    sa.first = (m == M);
    sa.rev = (M - m) & 1;
    sa.mom_in = mom[m & 1];
    sa.tile_ctr = dynamic_tiles ? tile_ctr + m : nullptr;
#ifdef PCF_TUNING
    sa.dbg = (m == dbg_date && grid <= 1024) ? g_sweep_dbg : nullptr;
#endif
    if (m < M && nccl_path) PCF_TRY(allreduce_sum(c, mom[m & 1], 8));
    sa.out = (m > 1) ? mom[(m - 1) & 1] : final_out(c);
    PeerLink l = c.link;
    l.host_err = c.perr_dev;
    l.call_first = c.call_first;
    l.call_last = c.call_last;
    l.gather = (m == 1) ? 1 : 0;  // the final sums: the publishing block also collects them (reduce.cuh)
    if (use_peer(c)) {
      l.seq = c.call_first + (unsigned long long)(M - m);  // published by this launch; it consumes seq - 1
    } else {
      l.world = 1;
      l.seq = 0;
    }
    if (w8) PCF_TRY((launch_sweep<uint8_t, Shape>(c, m == 1, pdl && m < M, grid, stages, sa, l)));
    else PCF_TRY((launch_sweep<uint16_t, Shape>(c, m == 1, pdl && m < M, grid, stages, sa, l)));
    c.launches++;
  }
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

// Host driver for one GPU. Enqueues everything on c.stream; the job-wide (sum, sumsq) of the discounted cash flows lands
// in final_out(c)[0..1] (NCCL path: this GPU's partial sums in c.d_out[0..1]); c.d_out[8..23] holds the per-date moment
// vectors of the per-date chain.
int run_mc_amer(Ctx& c, const pcf_params& p, Shard pairs, const double* d_replay, size_t ws_offset) {
  const int M = p.M;
  if (M > kMaxDates) {
    set_last_error("mc_amer: M exceeds 2048 exercise dates (include/pcf.h)");
    return PCF_EINVAL;
  }
  const long long H = pairs.size(), Nl = 2 * H;
  const double dt = p.T / M;
  int shape = kDefaultSweepShape, stages = kDefaultSweepStages;
  if (const char* v = tuning_env("PCF_AMER_SHAPE")) shape = atoi(v);   // <consumer warps><CTAs per SM>
  if (const char* v = tuning_env("PCF_AMER_SWEEP")) stages = std::max(2, std::min(kMaxStages, atoi(v)));
  const long long tile = 128LL * (shape / 10);
  if (tile < 128 || tile > kSweepMaxTile) {
    set_last_error("unknown PCF_AMER_SHAPE");
    return PCF_EINVAL;
  }
  const long long Np = (Nl + tile - 1) / tile * tile;  // padded row length: whole tiles of the sweep (128-byte multiples)
  char* base = (char*)c.workspace + ws_offset;
  double* paths = (double*)base;
  void* when = (void*)(paths + (size_t)M * Np);

  AmerArgs a;
  a.S0 = p.S0; a.E = p.E; a.cp = p.cp; a.M = M;
  a.adt = (p.r - 0.5 * p.sigma * p.sigma) * dt;
  a.cs = d_replay ? p.sigma : p.sigma * sqrt(dt);
  a.p0 = pairs.begin; a.H = H; a.Np = Np; a.seed = p.seed; a.w = d_replay;

  {
    static thread_local double tab[kDiscFwd + 2 * (kMaxDates + 1)];
    // path kernel: e^a 2^(+-j/32), a = (r - sigma^2/2) dt, in long double then rounded once
    const long double ea = expl((long double)a.adt);
    for (int j = 0; j < kExpEntries; ++j) {
      tab[2 * j] = (double)(ea * exp2l((long double)j / kExpEntries));
      tab[2 * j + 1] = (double)(ea * exp2l(-(long double)j / kExpEntries));
    }
    // discount tables, evaluated on the host with the reference's own expressions (glibc exp)
    for (int k = 0; k <= M; ++k) {
      tab[kDiscFwd + k] = exp(-p.r * dt * (double)k);            // exp(-r*dt*(exercise_when[n]-m))   mc_amer.cpp:50
      tab[kDiscFwd + (M + 1) + k] = exp(-p.r * (double)k * dt);  // exp(-r*exercise_when[n]*dt)       mc_amer.cpp:110
    }
    PCF_CUDA(cudaMemcpyToSymbolAsync(c_amer_tab, tab, sizeof(double) * (kDiscFwd + 2 * (M + 1)), 0,
                                     cudaMemcpyHostToDevice, c.stream));
  }
  const size_t gen_smem = kTableSmemBytes + (size_t)kExpEntries * kRep16 * sizeof(Pair);
  if (d_replay) {
    int grid_gen = grid_for(c, H, kAmerBlock, 2);
    amer_paths_kernel<true, 1, 2><<<grid_gen, kAmerBlock, gen_smem, c.stream>>>(a, c.d_tables, paths);
  } else {
    // launch shape: PCF_AMER_GEN = <pairs per thread><CTAs per SM> (PCF_TUNING builds)
    const char* v = tuning_env("PCF_AMER_GEN");
    const int variant = v ? atoi(v) : 32;  // ncu, 1e8 x 50: 32 -> 9.1 ms, 22 -> 9.6, 23 -> 9.9, 14 -> 10.5
#define PCF_GEN_CASE(P, B)                                                                       \
  case P * 10 + B: {                                                                             \
    int grid_gen = grid_for(c, (H + P - 1) / P, kAmerBlock, B);                                  \
    amer_paths_kernel<false, P, B><<<grid_gen, kAmerBlock, gen_smem, c.stream>>>(a, c.d_tables, paths); \
  } break;
    switch (variant) {
#ifdef PCF_TUNING
      PCF_GEN_CASE(1, 4)
      PCF_GEN_CASE(2, 2)
      PCF_GEN_CASE(2, 3)
      PCF_GEN_CASE(4, 1)
      PCF_GEN_CASE(4, 2)
      PCF_GEN_CASE(6, 1)
#endif
      PCF_GEN_CASE(3, 2)
      default:
        set_last_error("unknown PCF_AMER_GEN");
        return PCF_EINVAL;
    }
#undef PCF_GEN_CASE
  }
  c.launches++;
  if (Np != Nl) {
    amer_pad_kernel<<<(unsigned)std::min<long long>(((Np - Nl) * M + 255) / 256, 1024), 256, 0, c.stream>>>(paths, Nl, Np, M, p.cp);
    c.launches++;
  }
  PCF_CUDA(cudaGetLastError());

  // Backward sweep m = M .. 1 (mc_amer.cpp:23-27, 31, 109-111). The iteration of date m consumes the moments of date m
  // and produces those of date m-1 (date 1: the final sums).
  switch (shape) {
#ifdef PCF_TUNING
    case 83: return run_sweep_chain<SweepShape<8, 3>>(c, p, paths, when, Np, stages);
    case 122: return run_sweep_chain<SweepShape<12, 2>>(c, p, paths, when, Np, stages);
    case 161: return run_sweep_chain<SweepShape<16, 1>>(c, p, paths, when, Np, stages);
    case 201: return run_sweep_chain<SweepShape<20, 1>>(c, p, paths, when, Np, stages);
    case 281: return run_sweep_chain<SweepShape<28, 1>>(c, p, paths, when, Np, stages);
    case 112: return run_sweep_chain<SweepShape<11, 2>>(c, p, paths, when, Np, stages);
    case 73: return run_sweep_chain<SweepShape<7, 3>>(c, p, paths, when, Np, stages);
    case 241: return run_sweep_chain<SweepShape<24, 1>>(c, p, paths, when, Np, stages);
#endif
    case 231: return run_sweep_chain<SweepShape<23, 1>>(c, p, paths, when, Np, stages);
  }
  set_last_error("unknown PCF_AMER_SHAPE");
  return PCF_EINVAL;
}

#ifdef PCF_TUNING
// Tuning builds only: (smid, consumer cycles in the tile loop, tiles) of every CTA of one date's sweep kernel
extern "C" __attribute__((visibility("default"))) int pcf_debug_sweep(unsigned long long* out, int n) {
  if (!g_sweep_dbg) return PCF_ENOINIT;
  return cudaMemcpy(out, g_sweep_dbg, sizeof(unsigned long long) * (size_t)std::min(n, 3 * 1024), cudaMemcpyDeviceToHost) ==
                 cudaSuccess
             ? PCF_OK
             : PCF_ECUDA;
}
#endif

size_t amer_workspace_bytes(long long local_pairs, int M) {
  size_t Np = ((2 * (size_t)local_pairs + 15) & ~(size_t)15) + kSweepMaxTile;  // rows are padded to whole tiles
  // paths, exercise dates, one tile counter per date
  return (size_t)M * Np * 8 + amer_when_area((long long)Np, M) + sizeof(unsigned int) * (size_t)(M + 1) + 256;
}

}  // namespace pcf
