// amer_kernels.cu -- the reference's own early-exercise Monte Carlo scheme (a6 + a7 + a8).
//   reference include/common.h:168-208 (pathsfinder), src/mc_amer.cpp:23-113 (backward sweep),
//   include/common.h:98-141 (3x3 inverse + mat-vec).
//
// HBM layout (per GPU, n = local path index, Nl local paths = 2 * local antithetic pairs):
//   paths  double[M][Nl]   row m-1 holds S at date m (row 0 of the reference, S0, is never read by the
//                          sweep and is not stored). Pair p owns columns p and p + Nl/2, exactly the
//                          reference's antithetic halves; both stores of a warp are 256 B contiguous.
//   when   int32[Np]       exercise date; bit 30 set when the cash flow was booked by the regression
//                          branch (mc_amer.cpp:100-103), which books payoff(S - E, E) [SURVEY F1]
//   cash   double[Nl]      TRUE payoff at paths[when][n] -- the value the reference re-gathers at
//                          mc_amer.cpp:50; carrying it turns that row gather into a coalesced read.
// exercise_st of the reference is a pure function of (cash, flag): st = flag ? payoff(cp*cash, E)
// : cash, because cp*cash == S - E exactly for an in-the-money path.
//
// Rows and state arrays are padded to a multiple of 4 paths (Np) with never-in-the-money dummies so that every
// thread streams whole 32-byte sectors (4 paths) with 16-byte vector loads/stores.
//
// Backward sweep: ONE fused kernel per exercise date. amer_sweep_kernel for date m (a) waits for the moments of
// date m (its own, or -- multi-GPU -- every rank's, arriving in the NVLink mailbox), solves the 3x3 normal
// equations per block in the reference's operation order without FMA contraction, (b) applies the exercise
// decision of date m to the state held in registers, (c) accumulates the regression moments of date m-1 from
// that updated state and row m-1, and (d) its last block publishes them. Compared with a moments pass plus a
// decision pass this reads the 12 B/path state once per date instead of twice and halves the launches.
#include "common.cuh"
#include "reduce.cuh"
#include "rng.cuh"
#include <cstdlib>

namespace pcf {

constexpr int kAmerBlock = 256;
typedef int when_t;  // exercise date + flag. (uint16 was tried: 10% fewer bytes but 8% SLOWER sweeps -- 8-byte
                     // quarter-sector state stores; profiles/r1_notes.md)
constexpr int kQuirkBit = 1 << 30;
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
                                                                   double* __restrict__ paths,
                                                                   when_t* __restrict__ when,
                                                                   double* __restrict__ cash) {
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
    // mc_amer.cpp:23-27: exercise_when = M, exercise_st = payoff(S_M)
#pragma unroll
    for (int q = 0; q < kPairs; ++q) {
      when[pp[q]] = (when_t)a.M;
      when[pp[q] + a.H] = (when_t)a.M;
      cash[pp[q]] = payoff(Sp[q], a.E, a.cp);
      cash[pp[q] + a.H] = payoff(Sm[q], a.E, a.cp);
    }
  }
}

// Padding columns [2H, Np): S chosen so that payoff == 0 at every date, state = (M, 0).
__global__ void amer_pad_kernel(double* __restrict__ paths, when_t* __restrict__ when, double* __restrict__ cash,
                                long long Nl, long long Np, int M, int cp) {
  const long long pad = Np - Nl;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pad * (M + 1);
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / pad, col = Nl + i % pad;
    if (row < M) {
      paths[row * Np + col] = cp > 0 ? 0.0 : 1e300;  // call: S - E < 0;  put: E - S < 0
    } else {
      when[col] = (when_t)M;
      cash[col] = 0.0;
    }
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
//   kDecide:  apply the exercise decision of date m (rows S_m) from the moments of date m
//   kMoments: accumulate the moments of date m-1 (rows S_prev) and publish them
// out[0..7] = n_itm, Sx, Sx^2, Sx^3, Sx^4, Sy, Syx, Syx^2 with x = S - E, y = discounted cash flow; products
// are formed exactly like the reference forms them (left to right, no FMA): only the summation order differs.
constexpr int kMomFold = 4;
constexpr int kSweepUnroll = 2;
constexpr int kSweepBlock = 128;
constexpr int kSweepBlocksPerSM = 3;

__device__ __forceinline__ void amer_decide_one(double S, int& wq, double& cs, bool& changed, int mode, double c0,
                                                double c1, double c2, double E, int cp, int m,
                                                const double* s_disc) {
  const double pv = payoff(S, E, cp);
  if (!(pv > 0.0)) return;
  if (mode == 2 || mode == 3) {
    const double x = __dadd_rn(S, -E);
    if (mode == 2 && x == -1.0) return;  // the reference's sentinel collision (mc_amer.cpp:32,98)
    const double yhat = __dadd_rn(__dadd_rn(c0, __dmul_rn(c1, x)), __dmul_rn(c2, __dmul_rn(x, x)));
    // reference rule: payoff of the SHIFTED value (mc_amer.cpp:100); PCF_FLAG_AMER_LSM (mode 3): the true payoff
    const double pq = (mode == 2) ? payoff(x, E, cp) : pv;
    if (pq > yhat) {
      wq = (mode == 2) ? (m | kQuirkBit) : m;  // mode 3 books the true payoff: st == cash
      cs = pv;
      changed = true;
    }
  } else {
    // <= 2 paths in the money (mc_amer.cpp:75-83): true payoff against the discounted cash flow
    const double cont = __dmul_rn(s_disc[(wq & ~kQuirkBit) - m], cs);
    if (pv > cont) {
      wq = m;
      cs = pv;
      changed = true;
    }
  }
}

__device__ __forceinline__ void amer_moment_terms(double S, int wq, double cs, double E, int cp, int m,
                                                  const double* s_disc, double (&run)[8]) {
  if (payoff(S, E, cp) > 0.0) {
    const double ex = __dadd_rn(S, -E);
    const double cont = __dmul_rn(s_disc[(wq & ~kQuirkBit) - m], cs);
    const double ex2 = __dmul_rn(ex, ex), ex3 = __dmul_rn(ex2, ex), ex4 = __dmul_rn(ex3, ex);
    const double yx = __dmul_rn(cont, ex), yx2 = __dmul_rn(yx, ex);
    run[0] += 1.0;
    run[1] += ex;
    run[2] += ex2;
    run[3] += ex3;
    run[4] += ex4;
    run[5] += cont;
    run[6] += yx;
    run[7] += yx2;
  }
}

template <bool kDecide, bool kMoments>
__global__ void __launch_bounds__(kSweepBlock, kSweepBlocksPerSM) amer_sweep_kernel(
    const double* __restrict__ S_m, const double* __restrict__ S_prev, when_t* __restrict__ when,
    double* __restrict__ cash, long long Np, double E, int cp, int m, int M,
    const double* __restrict__ mom_in, PeerLink link_in, PeerLink link_out, double* partials,
    unsigned int* ticket, double* mom_out, int* err_flag, int lsm) {
  __shared__ double smem[8 * 2 * 32];
  __shared__ double s_mom[kXchgVals];
  __shared__ double s_coef[3];
  __shared__ int s_mode;  // 0 skip, 1 few-paths branch, 2 regression branch
  extern __shared__ double s_disc[];  // lanes index it with different k: shared memory, not constant
  for (int k = threadIdx.x; k <= M; k += blockDim.x) s_disc[k] = c_disc_fwd[k];
  int mode = 0;
  double c0 = 0.0, c1 = 0.0, c2 = 0.0;
  if (kDecide) {
    // moments of date m: from every rank's publication in this GPU's mailbox (multi-GPU), else local / all-reduced
    if (link_in.world > 1) {
      peer_gather<kXchgVals>(link_in, s_mom);
    } else {
      if (threadIdx.x < kXchgVals) s_mom[threadIdx.x] = mom_in[threadIdx.x];
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
          s_mode = lsm ? 3 : 2;
          s_coef[0] = coef[0]; s_coef[1] = coef[1]; s_coef[2] = coef[2];
        } else {
          s_mode = 0;
          if (blockIdx.x == 0) atomicExch(err_flag, PCF_ESINGULAR);  // common.h:115-117
        }
      }
    }
  }
  __syncthreads();
  if (kDecide) {
    mode = s_mode;
    c0 = s_coef[0]; c1 = s_coef[1]; c2 = s_coef[2];
  }
  Comp acc[8];
  double run[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int fold = 0;
  const long long quads = Np >> 2;
  const double2* Sm2 = reinterpret_cast<const double2*>(S_m);
  const double2* Sp2 = reinterpret_cast<const double2*>(S_prev);
  int4* W4 = reinterpret_cast<int4*>(when);
  double2* C2 = reinterpret_cast<double2*>(cash);
  const long long T = (long long)gridDim.x * blockDim.x;
  // kSweepUnroll quads (of 4 paths = one 32-byte sector per array) per thread iteration: all 7 x kSweepUnroll
  // 16-byte loads are issued before the first use, which keeps ~200 B per thread in flight -- the kernel is
  // DRAM-latency bound otherwise (ncu: long_scoreboard).
  for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < quads; base += T * kSweepUnroll) {
    double2 sa[kSweepUnroll], sb[kSweepUnroll], pa[kSweepUnroll], pb[kSweepUnroll], ca[kSweepUnroll], cb[kSweepUnroll];
    int4 w[kSweepUnroll];  // unpacked dates
    bool live[kSweepUnroll];
#pragma unroll
    for (int u = 0; u < kSweepUnroll; ++u) {
      const long long i = base + u * T;
      live[u] = i < quads;
      const long long j = live[u] ? i : base;  // clamp: loads stay in bounds, results are discarded
      sa[u] = sb[u] = pa[u] = pb[u] = make_double2(0, 0);
      if (kDecide) {
        sa[u] = __ldcs(Sm2 + 2 * j);
        sb[u] = __ldcs(Sm2 + 2 * j + 1);
      }
      if (kMoments) {
        pa[u] = __ldcs(Sp2 + 2 * j);
        pb[u] = __ldcs(Sp2 + 2 * j + 1);
      }
      w[u] = W4[j];
      ca[u] = C2[2 * j];
      cb[u] = C2[2 * j + 1];
    }
#pragma unroll
    for (int u = 0; u < kSweepUnroll; ++u) {
      if (!live[u]) continue;
      const long long i = base + u * T;
      if (kDecide && mode != 0) {
        bool changed = false;
        amer_decide_one(sa[u].x, w[u].x, ca[u].x, changed, mode, c0, c1, c2, E, cp, m, s_disc);
        amer_decide_one(sa[u].y, w[u].y, ca[u].y, changed, mode, c0, c1, c2, E, cp, m, s_disc);
        amer_decide_one(sb[u].x, w[u].z, cb[u].x, changed, mode, c0, c1, c2, E, cp, m, s_disc);
        amer_decide_one(sb[u].y, w[u].w, cb[u].y, changed, mode, c0, c1, c2, E, cp, m, s_disc);
        if (changed) {  // whole sectors back: no partial-sector fill from DRAM
          W4[i] = w[u];
          C2[2 * i] = ca[u];
          C2[2 * i + 1] = cb[u];
        }
      }
      if (kMoments) {
        amer_moment_terms(pa[u].x, w[u].x, ca[u].x, E, cp, m - 1, s_disc, run);
        amer_moment_terms(pa[u].y, w[u].y, ca[u].y, E, cp, m - 1, s_disc, run);
        amer_moment_terms(pb[u].x, w[u].z, cb[u].x, E, cp, m - 1, s_disc, run);
        amer_moment_terms(pb[u].y, w[u].w, cb[u].y, E, cp, m - 1, s_disc, run);
      }
    }
    if (kMoments && ++fold == kMomFold) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[k].add(run[k]);
        run[k] = 0.0;
      }
      fold = 0;
    }
  }
  if (kMoments) {
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k].add(run[k]);
    grid_reduce<8>(acc, smem, partials, ticket, mom_out, &link_out);
  }
}

// mc_amer.cpp:109-111: sum of discounted booked cash flows (+ sum of squares for the error bar).
__global__ void __launch_bounds__(kAmerBlock) amer_final_kernel(const when_t* __restrict__ when,
                                                                const double* __restrict__ cash,
                                                                long long Nl, double E, int cp, int M,
                                                                PeerLink link, double* partials,
                                                                unsigned int* ticket, double* out) {
  __shared__ double smem[2 * 2 * 32];
  extern __shared__ double s_disc[];
  for (int k = threadIdx.x; k <= M; k += blockDim.x) s_disc[k] = c_disc_abs[k];
  __syncthreads();
  BlockedComp<8> s1, s2;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < Nl;
       n += (long long)gridDim.x * blockDim.x) {
    int wq = when[n];
    double cs = cash[n];
    double st = (wq & kQuirkBit) ? payoff((double)cp * cs, E, cp) : cs;
    double v = (st != 0.0) ? __dmul_rn(s_disc[wq & ~kQuirkBit], st) : 0.0;
    s1.add(v);
    s2.add(v * v);
  }
  Comp v[2] = {s1.finish(), s2.finish()};
  grid_reduce<2>(v, smem, partials, ticket, out, &link);
}

// Host driver for one GPU. Enqueues everything on c.stream; result (sum, sumsq of discounted cash
// flows over local paths) lands in c.d_out[0..1]; c.d_out[8..15] is the per-date moment vector.
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
  double* cash = paths + (size_t)M * Np;
  when_t* when = (when_t*)(cash + Np);

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
    amer_paths_kernel<true, 1, 2><<<grid_gen, kAmerBlock, gen_smem, c.stream>>>(a, c.d_tables, paths, when, cash);
  } else {
    // launch shape: PCF_AMER_GEN = <pairs per thread><CTAs per SM> (tuning knob)
    const char* v = getenv("PCF_AMER_GEN");
    const int variant = v ? atoi(v) : 22;
#define PCF_GEN_CASE(P, B)                                                                                    \
  case P * 10 + B: {                                                                                          \
    int grid_gen = grid_for(c, (H + P - 1) / P, kAmerBlock, B);                                               \
    amer_paths_kernel<false, P, B><<<grid_gen, kAmerBlock, gen_smem, c.stream>>>(a, c.d_tables, paths, when, cash); \
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
    amer_pad_kernel<<<1, 128, 0, c.stream>>>(paths, when, cash, Nl, Np, M, p.cp);
    c.launches++;
  }
  PCF_CUDA(cudaGetLastError());

  // Backward sweep m = M-1 .. 1 (mc_amer.cpp:31). Kernel for date m consumes the moments of date m and produces
  // those of date m-1. Peer path: moments travel through the NVLink mailboxes (publish in the producing kernel,
  // gather in the consuming one); NCCL path: an all-reduce of the 8 doubles between two kernels.
  int grid = grid_for(c, (Np / 4 + kSweepUnroll - 1) / kSweepUnroll, kSweepBlock, kSweepBlocksPerSM);
  const int lsm = (p.flags & PCF_FLAG_AMER_LSM) ? 1 : 0;
  const size_t dsm = sizeof(double) * (M + 1);
  double* mom[2] = {c.d_out + 8, c.d_out + 16};
  auto row = [&](int m) { return paths + (size_t)(m - 1) * Np; };
  PeerLink none = c.link;
  none.world = 1;
  if (M >= 2) {
    PeerLink l_out = next_link(c);
    amer_sweep_kernel<false, true><<<grid, kSweepBlock, dsm, c.stream>>>(
        nullptr, row(M - 1), when, cash, Np, p.E, p.cp, M, M, nullptr, none, l_out, c.d_partials, c.d_ticket,
        mom[(M - 1) & 1], c.d_flag, lsm);
    c.launches++;
    for (int m = M - 1; m >= 1; --m) {
      const PeerLink l_in = l_out;
      if (!use_peer(c)) PCF_TRY(allreduce_sum(c, mom[m & 1], 8));
      if (m > 1) {
        l_out = next_link(c);
        amer_sweep_kernel<true, true><<<grid, kSweepBlock, dsm, c.stream>>>(
            row(m), row(m - 1), when, cash, Np, p.E, p.cp, m, M, mom[m & 1], l_in, l_out, c.d_partials,
            c.d_ticket, mom[(m - 1) & 1], c.d_flag, lsm);
      } else {
        amer_sweep_kernel<true, false><<<grid, kSweepBlock, dsm, c.stream>>>(
            row(1), nullptr, when, cash, Np, p.E, p.cp, 1, M, mom[1], l_in, none, c.d_partials, c.d_ticket,
            nullptr, c.d_flag, lsm);
      }
      c.launches++;
    }
  }
  *final_link = next_link(c);
  int grid_fin = grid_for(c, Np, kAmerBlock, 8);
  amer_final_kernel<<<grid_fin, kAmerBlock, dsm, c.stream>>>(when, cash, Np, p.E, p.cp, M, *final_link,
                                                            c.d_partials, c.d_ticket, c.d_out);
  c.launches++;
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

size_t amer_workspace_bytes(long long local_pairs, int M) {
  size_t Np = (2 * (size_t)local_pairs + 3) & ~(size_t)3;
  return (size_t)M * Np * 8 + Np * 8 + Np * 4 + 256;
}

}  // namespace pcf
