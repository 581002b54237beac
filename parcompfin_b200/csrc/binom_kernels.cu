// binom_kernels.cu -- the binomial-formula term sum (a9), log space, sm_100a FP64.
//   reference src/binom_embar.cpp:5-50 (binom) and include/common.h:63-72 (comb).
//
// The reference evaluates ln C(N,i) by re-summing up to N logarithms per term (O(N^2) overall,
// SURVEY F2) and accumulates ~1e-10 of rounding error by N = 1e5 (F4). Here every term pair
// (i, N-i) is one thread iteration and the log-weight is the saddle-point form (C. Loader, "Fast
// and accurate computation of binomial probabilities", 2000):
//   ln b(x;N,p) = d(N) - d(x) - d(N-x) - D(x;Np) - D(N-x;Nq) - 1/2 ln(2 pi x (N-x)/N)
//   d(n) = ln n! - [(n+1/2) ln n - n + 1/2 ln 2pi]  (Stirling error), D(x;m) = x ln(x/m) + m - x
// which has no large cancelling terms, so the O(N) sum is accurate to ~1e-13 at N = 1e8.
// u, d, p, q come from the host, derived with the reference's own expressions (F5).
#include <cmath>
#include "common.cuh"
#include "reduce.cuh"
#include "rng.cuh"
#include "binom_math.cuh"

namespace pcf {

constexpr int kBinomBlock = 256;
// profiles/r2_tune_binom_ways.log: with the screen four pairs side by side at 2 CTAs/SM (3.40 -> 3.00 ms at N = 2^31-1);
// without it the full routine wants occupancy, one pair at 3 CTAs/SM
constexpr int kDefaultBinomScreened = 42, kDefaultBinomUnscreened = 13;

template <int kScreenWays, int kMinBlocks>
__global__ void __launch_bounds__(kBinomBlock, kMinBlocks) binom_terms_kernel(BinomArgs a, const MathTables* __restrict__ tables,
                                                                           PeerLink link, double* partials,
                                                                           unsigned int* ticket, double* out) {
  __shared__ double smem[1 * 2 * 32];
  extern __shared__ __align__(16) unsigned char tab_smem[];
  const TableView tv = stage_tables(tables, tab_smem);
  Hoisted hc;
  hc.load();
  Comp acc;
  const long long T = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // pair 0 (terms 0 and N) has its own closed form; interior pairs i >= 1
  long long i = a.i0 + tid;
  if (i == 0) {
    acc.add(end_terms(a, tv));
    i += T;
  }
  if (a.screen) {
    // the loop carries x and N-x as doubles (exact below 2^53): no 64-bit integer conversions per pair
    double x = (double)i, nx = (double)(a.N - i);
    const double step = (double)T;
    // kScreenWays pairs are screened side by side: their logarithms are independent chains, and one pair per iteration
    // left the FP64 pipe waiting on its own results ("wait" was 25 % of the stall samples, profiles/r2g_ncu_binom_screen.txt).
    // The pairs that survive are still added in the order i, i+T, ...: the sum is bit-identical.
    for (; i + (kScreenWays - 1) * T < a.i1; i += kScreenWays * T, x += kScreenWays * step, nx -= kScreenWays * step) {
      bool dead[kScreenWays];
#pragma unroll
      for (int k = 0; k < kScreenWays; ++k) dead[k] = pair_dead(x + k * step, nx - k * step, a, tv, hc);
#pragma unroll
      for (int k = 0; k < kScreenWays; ++k)
        if (!dead[k]) acc.add(pair_terms(i + k * T, a, tv, hc));
    }
    for (; i < a.i1; i += T, x += step, nx -= step) {
      if (pair_dead(x, nx, a, tv, hc)) continue;  // both weights underflow: the pair adds exactly 0.0
      acc.add(pair_terms(i, a, tv, hc));
    }
  } else {
    for (; i < a.i1; i += T) acc.add(pair_terms(i, a, tv, hc));
  }
  if (a.add_mid && tid == 0)  // x == N-x: both halves of the pair are the same term (binom_embar.cpp:42-45)
    acc.add(0.5 * pair_terms(a.N / 2, a, tv, hc));
  Comp v[1] = {acc};
  grid_reduce<1>(v, smem, partials, ticket, out, &link);
}

// d(0..15), the Stirling-error table of stirlerr(): parameter-independent, uploaded once per context (pcf_api.cu ctx_open)
int upload_binom_tables(Ctx& c) {
  BinomArgs a;
  fill_binom_args(100.0, 100.0, 0.05, 0.2, 1.0, 1000, 1, a);
  PCF_CUDA(cudaMemcpyToSymbol(c_sfe, a.sfe, sizeof(a.sfe), 0, cudaMemcpyHostToDevice));
  return PCF_OK;
}

// `pairs` = this GPU's slice of i in [lo, until); result (partial undiscounted sum) -> final_out(c)[0].
int run_binom(Ctx& c, const pcf_params& p, Shard pairs, bool add_mid, const PeerLink& link) {
  BinomArgs a;
  fill_binom_args(p.S0, p.E, p.r, p.sigma, p.T, p.N, p.cp, a);
  a.i0 = pairs.begin; a.i1 = pairs.end; a.add_mid = add_mid ? 1 : 0;
  a.screen = (p.flags & PCF_FLAG_BINOM_NOSCREEN) ? 0 : 1;
  // launch shape: <pairs screened side by side><CTAs per SM> (PCF_BINOM_VARIANT in PCF_TUNING builds)
  const char* v = tuning_env("PCF_BINOM_VARIANT");
  const int variant = v ? atoi(v) : (a.screen ? kDefaultBinomScreened : kDefaultBinomUnscreened);
#define PCF_BINOM_CASE(W, B)                                                                                  \
  case W * 10 + B: {                                                                                          \
    const int grid = grid_for(c, pairs.size(), kBinomBlock, B);                                               \
    binom_terms_kernel<W, B><<<grid, kBinomBlock, kTableSmemBytes, c.stream>>>(a, c.d_tables, link, c.d_partials, \
                                                                             c.d_ticket, final_out(c));       \
  } break;
  switch (variant) {
#ifdef PCF_TUNING
    PCF_BINOM_CASE(1, 4)
    PCF_BINOM_CASE(2, 4)
    PCF_BINOM_CASE(2, 3)
    PCF_BINOM_CASE(4, 3)
    PCF_BINOM_CASE(4, 4)
    PCF_BINOM_CASE(8, 2)
#endif
    PCF_BINOM_CASE(4, 2)
    PCF_BINOM_CASE(1, 3)
    default:
      set_last_error("unknown PCF_BINOM_VARIANT");
      return PCF_EINVAL;
  }
#undef PCF_BINOM_CASE
  c.launches++;
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

}  // namespace pcf
