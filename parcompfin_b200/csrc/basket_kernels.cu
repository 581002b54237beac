// basket_kernels.cu -- basket (a4 + a5) Monte Carlo kernels, sm_100a FP64: the reference's equicorrelation
// basket (constant-column fast path) and the general triangular / full normal transform. Same structure as mc_kernels.cu:
// one path per thread iteration, Philox normals in registers, compensated (sum, sum^2) reduced thread -> warp -> block ->
// grid. No tensor cores: the d x d mat-vec per path lives in registers (136 of ~600 executed FP64 instructions at d = 16).
#include "common.cuh"
#include "reduce.cuh"
#include "rng.cuh"

namespace pcf {

constexpr int kBlock = 256;

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
  kernel<<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, c.d_ticket, final_out(c));
  return PCF_OK;
}

// Launch shape of the native kernel: PCF_BASKET_GEN = <paths per thread><CTAs per SM> (tuning knob): 13 | 12.
// Measured at d = 16 (profiles/r1_notes.md, r1s_tune_basket_general.log). Guarded body, 1e9 paths: 13 -> 82 ms, 22 -> 88,
// 41 -> 92, 21 -> 102. Guard-free body (kExact), 2e8 paths: 12 -> 12.2 ms, 22 -> 12.2, 21 -> 12.4, 13 -> 12.9, 11 -> 14.4
// (guarded 13: 16.4). Default 12 for the guard-free instantiation, 13 otherwise; only these two are built (every shape
// costs 20 instantiations). The replay flavour (parity path) is built once per dimension.
template <int D>
static int launch_basket(Ctx& c, const BasketArgs& a, long long paths, bool replay, bool full, const PeerLink& link) {
  // replay (parity path): always the full-matrix sweep -- the upper triangle of c_L holds zeros for a Cholesky factor and
  // fma(0, z, bt) == bt exactly, so one instantiation per dimension serves both branches of mvn.h:68-76
  if (replay) return basket_launch(c, mc_basket_kernel<D, true, true, 1, 3>, a, paths, 1, link);
  const bool exact = a.d == D && !tuning_env("PCF_BASKET_GUARDED");  // A/B knob: keep the per-column tests
#ifdef PCF_TUNING
  const char* e = tuning_env("PCF_BASKET_GEN");
  const int shape = e ? atoi(e) : (exact ? 12 : 13);  // one-block body: 1 path x 2 CTAs/SM (126 registers) is fastest
#define PCF_BG(P, B)                                                                                              \
  (exact ? (full ? basket_launch(c, mc_basket_kernel<D, false, true, P, B, true>, a, paths, P, link)              \
                 : basket_launch(c, mc_basket_kernel<D, false, false, P, B, true>, a, paths, P, link))            \
         : (full ? basket_launch(c, mc_basket_kernel<D, false, true, P, B>, a, paths, P, link)                    \
                 : basket_launch(c, mc_basket_kernel<D, false, false, P, B>, a, paths, P, link)))
  switch (shape) {
    case 13: return PCF_BG(1, 3);
    case 12: return PCF_BG(1, 2);
    // (11, 21 and 22 were measured too, profiles/r1s_tune_basket_general.log)
    default:
      set_last_error("unknown PCF_BASKET_GEN");
      return PCF_EINVAL;
  }
#undef PCF_BG
#else
  if (exact)
    return full ? basket_launch(c, mc_basket_kernel<D, false, true, 1, 2, true>, a, paths, 1, link)
                : basket_launch(c, mc_basket_kernel<D, false, false, 1, 2, true>, a, paths, 1, link);
  return full ? basket_launch(c, mc_basket_kernel<D, false, true, 1, 3>, a, paths, 1, link)
              : basket_launch(c, mc_basket_kernel<D, false, false, 1, 3>, a, paths, 1, link);
#endif
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
  bool equi = !rp && !spec && !(p.flags & PCF_FLAG_BASKET_GENERAL);
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
    const char* v = tuning_env("PCF_BASKET_VARIANT");  // <paths per thread><CTAs per SM> (PCF_TUNING builds)
    const int variant = v ? atoi(v) : 61;
#define PCF_BASKET_CASE(P, B)                                                                              \
  case P * 10 + B: {                                                                                       \
    int grid = grid_for(c, (paths.size() + P - 1) / P, kBlock, B);                                         \
    mc_basket_equi_kernel<P, B><<<grid, kBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, \
                                                                            c.d_ticket, final_out(c));          \
  } break;
    switch (variant) {
#ifdef PCF_TUNING
      PCF_BASKET_CASE(1, 4)
      PCF_BASKET_CASE(2, 2)
      PCF_BASKET_CASE(2, 3)
      PCF_BASKET_CASE(3, 2)
      PCF_BASKET_CASE(4, 1)
      PCF_BASKET_CASE(4, 2)
      PCF_BASKET_CASE(8, 1)
#endif
      PCF_BASKET_CASE(6, 1)
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

}  // namespace pcf
