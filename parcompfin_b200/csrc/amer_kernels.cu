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
// Rows and `when` are padded to a multiple of 4 paths (Np) with never-in-the-money dummies so that every
// thread streams whole 32-byte sectors (4 paths) with 16-byte vector loads.
//
// Backward sweep: ONE fused kernel per exercise date. amer_sweep_kernel for date m (a) waits for the moments of
// date m (its own, or -- multi-GPU -- every rank's, arriving in the NVLink mailbox), solves the 3x3 normal
// equations per block in the reference's operation order without FMA contraction, (b) applies the exercise
// decision of date m to the dates held in registers, (c) accumulates the regression moments of date m-1 from
// that updated state and row m-1, and (d) its last block publishes them. The kernel of date 1 accumulates the
// final discounted sum (mc_amer.cpp:109-111) instead of (c); the kernel of date M initialises the state
// (mc_amer.cpp:23-27) instead of (a)-(b).
#include "common.cuh"
#include "reduce.cuh"
#include "rng.cuh"
#include <cstdlib>

namespace pcf {

constexpr int kAmerBlock = 256;
constexpr int kMaxDates = 2048;  // discount tables: constant memory -> staged into shared memory per block

__constant__ double c_disc_fwd[kMaxDates + 1];  // exp(-r*dt*k)        as mc_amer.cpp:50 evaluates it
__constant__ double c_disc_abs[kMaxDates + 1];  // exp(-r*k*dt)        as mc_amer.cpp:110 evaluates it

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
// 32-entry table (c_amer_T, e^a folded in on the host), one LDS.128 per step: 15 FP64 for both exponentials.
__constant__ double c_amer_T[2 * kExpEntries];  // (e^a 2^(j/32), e^a 2^(-j/32)), j = 0..31

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
    s_T[i].x = c_amer_T[2 * (i / kRep16)];
    s_T[i].y = c_amer_T[2 * (i / kRep16) + 1];
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
      for (int m = 1; m <= a.M; m += 2) {
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
          if (m + 1 <= a.M) {
            amer_step(Sp[q], Sm[q], z1, cs, my_T);
            __stcs(paths + (size_t)m * Np + pp[q], Sp[q]);
            __stcs(paths + (size_t)m * Np + pp[q] + a.H, Sm[q]);
          }
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

// a7 (mc_amer.cpp:41-106), one fused kernel per date; see the file header.
//   kMoments: accumulate the moments of date m-1 (row S_prev) after the decision of date m and publish them
//   kFinal:   date 1 -- accumulate the discounted booked cash flows and their squares (mc_amer.cpp:109-111)
//   first:    date M -- no decision; the state is initialised to when = M (mc_amer.cpp:23-27)
// mom_out[0..7] = n_itm, Sx, Sx^2, Sx^3, Sx^4, Sy, Syx, Syx^2 with x = S - E, y = discounted cash flow; products
// are formed exactly like the reference forms them (left to right, no FMA): only the summation order differs.
constexpr int kMomFold = 8;
constexpr int kSweepMaxBlock = 256;

// Exercise dates of one quad of paths, packed: uint8 x 4 (one 32-bit word) or uint16 x 4 (one 64-bit word).
template <typename WT>
struct WhenQuad;
template <>
struct WhenQuad<uint8_t> {
  typedef uint32_t Vec;
  static constexpr int kFlag = 0x80, kMask = 0x7f;
  static __device__ __forceinline__ void unpack(Vec v, int (&w)[4]) {
    w[0] = v & 0xff; w[1] = (v >> 8) & 0xff; w[2] = (v >> 16) & 0xff; w[3] = v >> 24;
  }
  static __device__ __forceinline__ Vec pack(const int (&w)[4]) {
    return (uint32_t)w[0] | ((uint32_t)w[1] << 8) | ((uint32_t)w[2] << 16) | ((uint32_t)w[3] << 24);
  }
};
template <>
struct WhenQuad<uint16_t> {
  typedef uint2 Vec;
  static constexpr int kFlag = 0x8000, kMask = 0x7fff;
  static __device__ __forceinline__ void unpack(Vec v, int (&w)[4]) {
    w[0] = v.x & 0xffff; w[1] = v.x >> 16; w[2] = v.y & 0xffff; w[3] = v.y >> 16;
  }
  static __device__ __forceinline__ Vec pack(const int (&w)[4]) {
    return make_uint2((uint32_t)w[0] | ((uint32_t)w[1] << 16), (uint32_t)w[2] | ((uint32_t)w[3] << 16));
  }
};

struct SweepArgs {
  const double* paths;  // row m-1 = date m, stride Np
  void* when;
  long long Np;
  double E;
  int cp, m, M;
  int first;            // date M: no decision, state := M
  int lsm;              // PCF_FLAG_AMER_LSM
  const double* mom_in;
  double* partials;
  unsigned int* ticket;
  double* out;          // kMoments: the 8 moments of date m-1; kFinal: sum, sumsq
  int* err_flag;
  int dbg;              // PCF_AMER_DBG (timing experiments only; results are wrong when set)
};

template <typename WT, bool kMoments, bool kFinal, int kUnroll>
__global__ void __launch_bounds__(kSweepMaxBlock) amer_sweep_kernel(SweepArgs a, PeerLink link_in, PeerLink link_out) {
  typedef WhenQuad<WT> WQ;
  typedef typename WQ::Vec WVec;
  constexpr int kFlag = WQ::kFlag, kMask = WQ::kMask;
  __shared__ double smem[8 * 2 * 32];
  __shared__ double s_mom[kXchgVals];
  __shared__ double s_coef[3];
  __shared__ int s_mode;  // 0 skip, 1 few-paths branch, 2 regression branch, 3 regression branch (LSM rule)
  extern __shared__ double s_disc[];  // lanes index it with different k: shared memory, not constant
  double* s_abs = s_disc + (a.M + 1);
  for (int k = threadIdx.x; k <= a.M; k += blockDim.x) {
    s_disc[k] = c_disc_fwd[k];
    if (kFinal) s_abs[k] = c_disc_abs[k];
  }
  const int m = a.m, cp = a.cp;
  const double E = a.E;
  if (!a.first) {
    // moments of date m: from every rank's publication in this GPU's mailbox (multi-GPU), else local / all-reduced
    if (link_in.world > 1) {
      peer_gather<kXchgVals>(link_in, s_mom);
    } else {
      if (threadIdx.x < kXchgVals) s_mom[threadIdx.x] = a.mom_in[threadIdx.x];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const double cnt = s_mom[0];
      if (cnt == 0.0) {
        s_mode = 0;                       // mc_amer.cpp:73
      } else if (cnt <= 2.0) {
        s_mode = 1;                       // mc_amer.cpp:75
      } else {
        double coef[3];
        if (solve3_reference_order(s_mom, coef)) {
          s_mode = a.lsm ? 3 : 2;
          s_coef[0] = coef[0]; s_coef[1] = coef[1]; s_coef[2] = coef[2];
        } else {
          s_mode = 0;
          if (blockIdx.x == 0) atomicExch(a.err_flag, PCF_ESINGULAR);  // common.h:115-117
        }
      }
    }
  } else if (threadIdx.x == 0) {
    s_mode = 0;
  }
  __syncthreads();
  const int mode = s_mode;
  const double c0 = s_coef[0], c1 = s_coef[1], c2 = s_coef[2];
  const bool first = a.first != 0;

  // Everything below works on cx = cp*(S - E) = fma(sgn, S, -sgn*E): bit-identical to the reference's
  // (double)cp*(S - E) (negation is exact and rounding is symmetric), one DFMA. The regressor x = S - E is
  // sgn*cx, so x^2, x^4 and y x^2 are sign-free and Sx, Sx^3, Syx are sgn times the sums formed from cx.
  const double sgn = (double)cp, nE = -sgn * E;
  const double c1s = c1 * sgn;  // c1*x == (c1*sgn)*cx exactly
  // per-thread compensated totals live in shared memory (touched once per kMomFold iterations): 32 registers
  // less per thread, i.e. more resident warps for a kernel whose limiter is DRAM latency
  __shared__ double2 s_acc[8][kSweepMaxBlock];
#pragma unroll
  for (int k = 0; k < 8; ++k) s_acc[k][threadIdx.x] = make_double2(0.0, 0.0);
  double run[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int cnt = 0, fold = 0;
  auto fold_runs = [&]() {
    run[0] += (double)cnt;
    cnt = 0;
#pragma unroll
    for (int k = 0; k < (kFinal ? 2 : 8); ++k) {
      const double2 t = s_acc[k][threadIdx.x];
      Comp c(t.x, t.y);
      c.add(run[k]);
      s_acc[k][threadIdx.x] = make_double2(c.hi, c.lo);
      run[k] = 0.0;
    }
  };
  const long long Np = a.Np, quads = Np >> 2;
  const double* __restrict__ paths = a.paths;
  const double2* Sm2 = reinterpret_cast<const double2*>(paths + (size_t)(m - 1) * Np);
  const double2* Sp2 = reinterpret_cast<const double2*>(paths + (size_t)(kMoments ? m - 2 : 0) * Np);
  WVec* W = reinterpret_cast<WVec*>(a.when);
  const int booked = (mode == 2) ? (m | kFlag) : m;  // mode 3 (LSM rule) books the true payoff

  const int lane = threadIdx.x & 31;
  const long long T = (long long)gridDim.x * blockDim.x;
  // kUnroll quads (4 paths = one 32-byte sector per row) per thread iteration; every streaming load is issued
  // before the first use. Loop bounds are warp-uniform (the store vote below needs the whole warp).
  for (long long wb = (long long)blockIdx.x * blockDim.x + (threadIdx.x - lane); wb < quads; wb += T * kUnroll) {
    double2 sa[kUnroll], sb[kUnroll], pa[kUnroll], pb[kUnroll];
    WVec wv[kUnroll];
    bool live[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = wb + lane + u * T;
      live[u] = i < quads;
      const long long j = live[u] ? i : wb;  // clamp: loads stay in bounds, results are discarded
      sa[u] = __ldcs(Sm2 + 2 * j);
      sb[u] = __ldcs(Sm2 + 2 * j + 1);
      pa[u] = pb[u] = make_double2(0.0, 0.0);
      if (kMoments) {
        pa[u] = __ldcs(Sp2 + 2 * j);
        pb[u] = __ldcs(Sp2 + 2 * j + 1);
      }
      if (!first) wv[u] = W[j];
    }
    // (b) decisions of date m for every quad of this iteration; `src` ends up holding, per path, the spot whose
    // payoff is the path's cash flow: S_m when the path exercises at m, else paths[when][n] (gathered below)
    int w[kUnroll][4];
    double src[kUnroll][4];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = wb + lane + u * T;
      if (first) {
        w[u][0] = w[u][1] = w[u][2] = w[u][3] = a.M;
      } else {
        WQ::unpack(wv[u], w[u]);
      }
      src[u][0] = sa[u].x; src[u][1] = sa[u].y; src[u][2] = sb[u].x; src[u][3] = sb[u].y;
      bool changed = false;
      if (mode >= 2) {
        // regression branch, branch-free: every lane evaluates the fit, the update is a select
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double cx = fma(sgn, src[u][e], nE);  // cp*(S - E); payoff(S) = max(cx, 0)
          const double yhat = __dadd_rn(__dadd_rn(c0, __dmul_rn(c1s, cx)), __dmul_rn(c2, __dmul_rn(cx, cx)));
          // reference rule: payoff of the SHIFTED value, payoff(x, E) = max(cp*(x - E), 0) = max(cx - cp*E, 0)
          // (mc_amer.cpp:100); PCF_FLAG_AMER_LSM (mode 3): the true payoff, cx
          const double pq = (mode == 2) ? __dadd_rn(cx, nE) : cx;  // before the max(., 0): max(t,0) > y <=> t > y || 0 > y
          // x == -1 is the reference's sentinel collision (mc_amer.cpp:32,98): such a path is skipped by its pass 2
          const bool ex = live[u] && cx > 0.0 && !(mode == 2 && cx == -sgn) && (pq > yhat || 0.0 > yhat);
          w[u][e] = ex ? booked : w[u][e];
          changed |= ex;
        }
      } else if (mode == 1 && live[u]) {
        // <= 2 paths in the money (mc_amer.cpp:75-83): true payoff against the discounted cash flow
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double pv = payoff(src[u][e], E, cp);
          if (!(pv > 0.0)) continue;
          const int d = w[u][e] & kMask;
          const double cont = __dmul_rn(s_disc[d - m], payoff(__ldg(paths + (size_t)(d - 1) * Np + 4 * i + e), E, cp));
          if (pv > cont) {
            w[u][e] = m;
            changed = true;
          }
        }
      }
      // whole 128-byte lines back: the warp stores when any of its lanes changed (or initialises at date M)
      if (__any_sync(0xffffffffu, changed) || first) {
        if (live[u] && !(a.dbg & 2)) W[i] = WQ::pack(w[u]);
      }
    }
    // (c) gathers of paths[when][n] (mc_amer.cpp:50) for paths whose exercise date is older than m and whose cash
    // flow is needed: all issued before the first use, so a quad costs one more DRAM round trip, not four
    double cxp[kUnroll][4];  // cp*(S_{m-1} - E), zeroed when the path is out of the money at m-1
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = wb + lane + u * T;
      const double Sp[4] = {pa[u].x, pa[u].y, pb[u].x, pb[u].y};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = w[u][e] & kMask;
        bool need = live[u];
        if (kMoments) {
          const double c = fma(sgn, Sp[e], nE);
          need = need && c > 0.0;
          cxp[u][e] = need ? c : 0.0;
        }
        if (need && d != m && !(a.dbg & 1)) src[u][e] = __ldg(paths + (size_t)(d - 1) * Np + 4 * i + e);
        if (kFinal && !need) w[u][e] = 0;  // date 0: discount slot, never booked (s_abs[0] * 0)
      }
    }
    // (d) moments of date m-1 / final sum. Out-of-the-money (and dead) lanes add exact zeros.
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = w[u][e] & kMask;
        const double cr = fma(sgn, src[u][e], nE);
        const double cs = cr > 0.0 ? cr : 0.0;  // payoff(paths[when][n])
        if (kMoments) {
          const double ex = cxp[u][e];
          const bool in = ex > 0.0;
          const double cont = in ? __dmul_rn(s_disc[in ? d - (m - 1) : 0], cs) : 0.0;
          const double ex2 = __dmul_rn(ex, ex), ex3 = __dmul_rn(ex2, ex), ex4 = __dmul_rn(ex3, ex);
          const double yx = __dmul_rn(cont, ex), yx2 = __dmul_rn(yx, ex);
          cnt += in ? 1 : 0;
          run[1] += ex;
          run[2] += ex2;
          run[3] += ex3;
          run[4] += ex4;
          run[5] += cont;
          run[6] += yx;
          run[7] += yx2;
        }
        if (kFinal) {
          // exercise_st (mc_amer.cpp:103): the regression branch booked payoff(x, E) = max(cp*(x - E), 0), x = S - E
          const double sq = __dadd_rn(cs, nE);
          const double st = (w[u][e] & kFlag) ? (sq > 0.0 ? sq : 0.0) : cs;
          const double v = (d != 0 && st != 0.0) ? __dmul_rn(s_abs[d], st) : 0.0;
          run[0] += v;
          run[1] += v * v;
        }
      }
    }
    if (++fold == kMomFold) {
      fold_runs();
      fold = 0;
    }
  }
  fold_runs();
  Comp acc[8];
#pragma unroll
  for (int k = 0; k < (kFinal ? 2 : 8); ++k) {
    const double2 t = s_acc[k][threadIdx.x];
    acc[k] = Comp(t.x, t.y);
  }
  if (kMoments) {
    // back from cx-space to the reference's x = S - E: odd powers of x carry the sign of cp
    acc[1].hi *= sgn; acc[1].lo *= sgn;
    acc[3].hi *= sgn; acc[3].lo *= sgn;
    acc[6].hi *= sgn; acc[6].lo *= sgn;
    grid_reduce<8>(acc, smem, a.partials, a.ticket, a.out, &link_out);
  }
  if (kFinal) {
    Comp v[2];
    v[0] = acc[0]; v[1] = acc[1];
    grid_reduce<2>(v, smem, a.partials, a.ticket, a.out, &link_out);
  }
}

template <typename WT, bool kMoments, bool kFinal>
static void launch_sweep(int unroll, int grid, int block, size_t smem, cudaStream_t st, const SweepArgs& a,
                         const PeerLink& li, const PeerLink& lo) {
  switch (unroll) {
    case 1: amer_sweep_kernel<WT, kMoments, kFinal, 1><<<grid, block, smem, st>>>(a, li, lo); break;
    case 2: amer_sweep_kernel<WT, kMoments, kFinal, 2><<<grid, block, smem, st>>>(a, li, lo); break;
    default: amer_sweep_kernel<WT, kMoments, kFinal, 4><<<grid, block, smem, st>>>(a, li, lo); break;
  }
}

static inline size_t amer_when_bytes(int M) { return M <= WhenQuad<uint8_t>::kMask ? 1 : 2; }

// Host driver for one GPU. Enqueues everything on c.stream; result (sum, sumsq of discounted cash
// flows over local paths) lands in c.d_out[0..1]; c.d_out[8..23] holds the per-date moment vectors.
int run_mc_amer(Ctx& c, const pcf_params& p, Shard pairs, const double* d_replay, size_t ws_offset,
                PeerLink* final_link) {
  const int M = p.M;
  if (M > kMaxDates) {
    set_last_error("mc_amer: M exceeds kMaxDates");
    return PCF_EINVAL;
  }
  const long long H = pairs.size(), Nl = 2 * H;
  const double dt = p.T / M;
  // discount tables, evaluated on the host with the reference's own expressions (glibc exp)
  static thread_local double fwd[kMaxDates + 1], ab[kMaxDates + 1];
  for (int k = 0; k <= M; ++k) {
    fwd[k] = exp(-p.r * dt * (double)k);  // exp(-r*dt*(exercise_when[n]-m))   mc_amer.cpp:50
    ab[k] = exp(-p.r * (double)k * dt);   // exp(-r*exercise_when[n]*dt)       mc_amer.cpp:110
  }
  PCF_CUDA(cudaMemcpyToSymbolAsync(c_disc_fwd, fwd, sizeof(double) * (M + 1), 0, cudaMemcpyHostToDevice, c.stream));
  PCF_CUDA(cudaMemcpyToSymbolAsync(c_disc_abs, ab, sizeof(double) * (M + 1), 0, cudaMemcpyHostToDevice, c.stream));

  const long long Np = (Nl + 3) & ~3LL;  // padded row length
  char* base = (char*)c.workspace + ws_offset;
  double* paths = (double*)base;
  void* when = (void*)(paths + (size_t)M * Np);

  AmerArgs a;
  a.S0 = p.S0; a.E = p.E; a.cp = p.cp; a.M = M;
  a.adt = (p.r - 0.5 * p.sigma * p.sigma) * dt;
  a.cs = d_replay ? p.sigma : p.sigma * sqrt(dt);
  a.p0 = pairs.begin; a.H = H; a.Np = Np; a.seed = p.seed; a.w = d_replay;

  {
    // per-call table: e^a 2^(+-j/32), a = (r - sigma^2/2) dt, in long double then rounded once
    double Tt[2 * kExpEntries];
    const long double ea = expl((long double)a.adt);
    for (int j = 0; j < kExpEntries; ++j) {
      Tt[2 * j] = (double)(ea * exp2l((long double)j / kExpEntries));
      Tt[2 * j + 1] = (double)(ea * exp2l(-(long double)j / kExpEntries));
    }
    PCF_CUDA(cudaMemcpyToSymbolAsync(c_amer_T, Tt, sizeof(Tt), 0, cudaMemcpyHostToDevice, c.stream));
  }
  const size_t gen_smem = kTableSmemBytes + (size_t)kExpEntries * kRep16 * sizeof(Pair);
  if (d_replay) {
    int grid_gen = grid_for(c, H, kAmerBlock, 2);
    amer_paths_kernel<true, 1, 2><<<grid_gen, kAmerBlock, gen_smem, c.stream>>>(a, c.d_tables, paths);
  } else {
    // launch shape: PCF_AMER_GEN = <pairs per thread><CTAs per SM> (tuning knob)
    const char* v = getenv("PCF_AMER_GEN");
    const int variant = v ? atoi(v) : 22;
#define PCF_GEN_CASE(P, B)                                                                       \
  case P * 10 + B: {                                                                             \
    int grid_gen = grid_for(c, (H + P - 1) / P, kAmerBlock, B);                                  \
    amer_paths_kernel<false, P, B><<<grid_gen, kAmerBlock, gen_smem, c.stream>>>(a, c.d_tables, paths); \
  } break;
    switch (variant) {
      PCF_GEN_CASE(1, 4)
      PCF_GEN_CASE(2, 2)
      PCF_GEN_CASE(2, 3)
      PCF_GEN_CASE(3, 2)
      PCF_GEN_CASE(4, 1)
      PCF_GEN_CASE(4, 2)
      PCF_GEN_CASE(6, 1)
      default:
        set_last_error("unknown PCF_AMER_GEN");
        return PCF_EINVAL;
    }
#undef PCF_GEN_CASE
  }
  c.launches++;
  if (Np != Nl) {
    amer_pad_kernel<<<1, 128, 0, c.stream>>>(paths, Nl, Np, M, p.cp);
    c.launches++;
  }
  PCF_CUDA(cudaGetLastError());

  // Backward sweep m = M .. 1 (mc_amer.cpp:23-27, 31, 109-111). Kernel for date m consumes the moments of date m
  // and produces those of date m-1. Peer path: moments travel through the NVLink mailboxes (publish in the
  // producing kernel, gather in the consuming one); NCCL path: an all-reduce of the 8 doubles between two kernels.
  // Launch shape: PCF_AMER_SWEEP = "<quads per thread>,<threads per CTA>,<CTAs per SM>" (tuning knob).
  int unroll = 4, block = 128, per_sm = 4;
  if (const char* v = getenv("PCF_AMER_SWEEP")) {
    if (sscanf(v, "%d,%d,%d", &unroll, &block, &per_sm) != 3 || (unroll != 1 && unroll != 2 && unroll != 4) ||
        block < 32 || block > kSweepMaxBlock || block % 32 != 0 || per_sm < 1) {
      set_last_error("bad PCF_AMER_SWEEP");
      return PCF_EINVAL;
    }
  }
  const int grid = grid_for(c, (Np / 4 + unroll - 1) / unroll, block, per_sm);
  const bool w8 = amer_when_bytes(M) == 1;
  SweepArgs sa;
  sa.paths = paths; sa.when = when; sa.Np = Np; sa.E = p.E; sa.cp = p.cp; sa.M = M;
  sa.lsm = (p.flags & PCF_FLAG_AMER_LSM) ? 1 : 0;
  sa.dbg = getenv("PCF_AMER_DBG") ? atoi(getenv("PCF_AMER_DBG")) : 0;
  sa.partials = c.d_partials; sa.ticket = c.d_ticket; sa.err_flag = c.d_flag;
  double* mom[2] = {c.d_out + 8, c.d_out + 16};
  PeerLink none = c.link;
  none.world = 1;
  PeerLink l_in = none;
  for (int m = M; m >= 1; --m) {
    sa.m = m;
    sa.first = (m == M);
    sa.mom_in = mom[m & 1];
    if (m < M && !use_peer(c)) PCF_TRY(allreduce_sum(c, mom[m & 1], 8));
    const PeerLink l_out = next_link(c);
    if (m > 1) {
      sa.out = mom[(m - 1) & 1];
      const size_t dsm = sizeof(double) * (M + 1);
      if (w8) launch_sweep<uint8_t, true, false>(unroll, grid, block, dsm, c.stream, sa, l_in, l_out);
      else launch_sweep<uint16_t, true, false>(unroll, grid, block, dsm, c.stream, sa, l_in, l_out);
    } else {
      sa.out = c.d_out;
      const size_t dsm = 2 * sizeof(double) * (M + 1);
      if (w8) launch_sweep<uint8_t, false, true>(unroll, grid, block, dsm, c.stream, sa, l_in, l_out);
      else launch_sweep<uint16_t, false, true>(unroll, grid, block, dsm, c.stream, sa, l_in, l_out);
      *final_link = l_out;
    }
    l_in = l_out;
    c.launches++;
  }
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

size_t amer_workspace_bytes(long long local_pairs, int M) {
  size_t Np = (2 * (size_t)local_pairs + 3) & ~(size_t)3;
  return (size_t)M * Np * 8 + Np * amer_when_bytes(M) + 256;
}

}  // namespace pcf
