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

namespace pcf {

constexpr int kBinomBlock = 256;

struct BinomArgs {
  double S0, E;
  int cp;
  double Nd;                 // N as double
  long long N;
  long long i0, i1;          // this GPU's pair range, i in [i0, i1), i < ceil(N/2)
  int add_mid;               // this GPU also adds the middle term N/2 (N even), as binom_embar.cpp:42-45
  double np_hi, np_lo, nq_hi, nq_lo;   // N*p, N*q as double-double
  double lnp_hi, lnp_lo, lnq_hi, lnq_lo;
  double lnu_hi, lnu_lo, lnd_hi, lnd_lo;
  double stirl_N;            // d(N)
  double corr;               // N*(p+q-1): q = fl(1-p) is not exactly 1-p, and the sum is defined on
                             // the reference's (p, q) doubles (binom_embar.cpp:24-27)
  double sfe[16];            // d(0..15), host long-double values
};

__device__ __forceinline__ double stirlerr(double n, const BinomArgs& a) {
  if (n < 16.0) return a.sfe[(int)n];
  const double r = 1.0 / n, r2 = r * r;
  // 1/12 - 1/360 r^2 + 1/1260 r^4 - 1/1680 r^6 + 1/1188 r^8 - 691/360360 r^10
  double s = -691.0 / 360360.0;
  s = fma(s, r2, 1.0 / 1188.0);
  s = fma(s, r2, -1.0 / 1680.0);
  s = fma(s, r2, 1.0 / 1260.0);
  s = fma(s, r2, -1.0 / 360.0);
  s = fma(s, r2, 1.0 / 12.0);
  return s * r;
}

// deviance D(x; m) with m = m_hi + m_lo
__device__ __forceinline__ double bd0(double x, double m_hi, double m_lo) {
  const double diff = (x - m_hi) - m_lo;
  const double sum = x + m_hi;
  if (fabs(diff) < 0.1 * sum) {
    const double v = diff / sum, v2 = v * v;
    double s = diff * v;
    double ej = 2.0 * x * v;
#pragma unroll
    for (int j = 1; j <= 9; ++j) {
      ej *= v2;
      s = fma(ej, 1.0 / (double)(2 * j + 1), s);
    }
    return s;
  }
  return fma(x, log(x / m_hi), -diff);
}

// ln b(x; N, p) for 0 < x < N; lf = ln(2 pi x (N-x)/N) shared by the pair
__device__ __forceinline__ double lpmf_inner(double x, double nx, double sx, double snx, double lf,
                                             const BinomArgs& a) {
  double lc = a.stirl_N - sx - snx - bd0(x, a.np_hi, a.np_lo) - bd0(nx, a.nq_hi, a.nq_lo);
  return fma(-0.5, lf, lc) + a.corr;
}

// ln(S0 u^x d^(N-x)) - ln S0 = x ln u + (N-x) ln d, double-double products and sum
__device__ __forceinline__ double log_growth(double x, double nx, const BinomArgs& a) {
  double t1 = x * a.lnu_hi, e1 = fma(x, a.lnu_hi, -t1);
  double t2 = nx * a.lnd_hi, e2 = fma(nx, a.lnd_hi, -t2);
  double s = t1 + t2, bb = s - t1;
  double err = (t1 - (s - bb)) + (t2 - bb);
  double lo = err + e1 + e2 + fma(x, a.lnu_lo, nx * a.lnd_lo);
  return s + lo;
}

__device__ __forceinline__ double term(double lw, double x, double nx, const BinomArgs& a) {
  double w = exp(lw);
  if (!(w > 0.0)) return 0.0;  // weight underflow: the term is 0, not 0*inf
  double S = a.S0 * exp(log_growth(x, nx, a));
  return w * payoff(S, a.E, a.cp);
}

__device__ __forceinline__ double pair_terms(long long i, const BinomArgs& a) {
  const double x = (double)i, nx = (double)(a.N - i);
  double lw1, lw2;  // weights of "x ups" and of "N-x ups"
  if (i == 0) {
    lw1 = fma(a.Nd, a.lnq_hi, a.Nd * a.lnq_lo);  // q^N
    lw2 = fma(a.Nd, a.lnp_hi, a.Nd * a.lnp_lo);  // p^N
  } else {
    const double sx = stirlerr(x, a), snx = stirlerr(nx, a);
    const double lf = log(6.283185307179586476925286766559 * (x * (nx / a.Nd)));
    lw1 = lpmf_inner(x, nx, sx, snx, lf, a);
    lw2 = lpmf_inner(nx, x, snx, sx, lf, a);
  }
  return term(lw1, x, nx, a) + term(lw2, nx, x, a);
}

__global__ void __launch_bounds__(kBinomBlock) binom_terms_kernel(BinomArgs a, PeerLink link, double* partials,
                                                                  unsigned int* ticket, double* out) {
  __shared__ double smem[1 * 2 * 32];
  Comp acc;
  for (long long i = a.i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.i1;
       i += (long long)gridDim.x * blockDim.x)
    acc.add(pair_terms(i, a));
  if (a.add_mid && blockIdx.x == 0 && threadIdx.x == 0) {
    const double h = (double)(a.N / 2);
    const double sh = stirlerr(h, a);
    const double lf = log(6.283185307179586476925286766559 * (h * (h / a.Nd)));
    acc.add(term(lpmf_inner(h, h, sh, sh, lf, a), h, h, a));
  }
  Comp v[1] = {acc};
  grid_reduce<1>(v, smem, partials, ticket, out, &link);
}

static long double stirlerr_host(long double n) {
  if (n == 0) return 0.0L;
  const long double half_ln_2pi = 0.918938533204672741780329736405617639L;
  return lgammal(n + 1.0L) - ((n + 0.5L) * logl(n) - n + half_ln_2pi);
}

static void split(long double v, double& hi, double& lo) {
  hi = (double)v;
  lo = (double)(v - (long double)hi);
}

// Reference lattice parameters, binom_embar.cpp:19-27, evaluated with the same libm calls in the
// same order (the sqrt(beta^2-1) cancellation makes any algebraic rewrite visible).
void binom_lattice(double r, double sigma, double T, long long N, double& u, double& d, double& p,
                   double& q) {
  double dt = (double)T / (double)N;
  double beta = 0.5 * (exp(-r * dt) + exp((r + sigma * sigma) * dt));
  u = beta + sqrt(beta * beta - 1);
  d = beta - sqrt(beta * beta - 1);
  double R = exp(r * dt);
  p = (R - d) / (u - d);
  q = 1 - p;
}

// `pairs` = this GPU's slice of i in [lo, until); result (partial undiscounted sum) -> c.d_out[0].
int run_binom(Ctx& c, const pcf_params& p, Shard pairs, bool add_mid, const PeerLink& link) {
  BinomArgs a;
  double u, d, pp, q;
  binom_lattice(p.r, p.sigma, p.T, p.N, u, d, pp, q);
  a.S0 = p.S0; a.E = p.E; a.cp = p.cp; a.N = p.N; a.Nd = (double)p.N;
  a.i0 = pairs.begin; a.i1 = pairs.end; a.add_mid = add_mid ? 1 : 0;
  a.np_hi = a.Nd * pp; a.np_lo = fma(a.Nd, pp, -a.np_hi);  // exact double-double products
  a.nq_hi = a.Nd * q;  a.nq_lo = fma(a.Nd, q, -a.nq_hi);
  a.corr = (double)((long double)p.N * (((long double)pp + (long double)q) - 1.0L));
  split(logl((long double)pp), a.lnp_hi, a.lnp_lo);
  split(logl((long double)q), a.lnq_hi, a.lnq_lo);
  split(logl((long double)u), a.lnu_hi, a.lnu_lo);
  split(logl((long double)d), a.lnd_hi, a.lnd_lo);
  if (p.N < 64) a.stirl_N = (double)stirlerr_host((long double)p.N);
  else {
    long double n = (long double)p.N, r2 = 1.0L / (n * n);
    a.stirl_N = (double)((1.0L / 12 - (1.0L / 360 - (1.0L / 1260 - (1.0L / 1680 - (1.0L / 1188) * r2) * r2) * r2) * r2) / n);
  }
  for (int k = 0; k < 16; ++k) a.sfe[k] = (double)stirlerr_host((long double)k);
  int grid = grid_for(c, pairs.size(), kBinomBlock, 8);
  binom_terms_kernel<<<grid, kBinomBlock, 0, c.stream>>>(a, link, c.d_partials, c.d_ticket, c.d_out);
  c.launches++;
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

}  // namespace pcf
