// binom_math.cuh -- one term pair (i, N-i) of the binomial-formula sum, host/device (a9).
//   reference src/binom_embar.cpp:34-46 (the pairing) and include/common.h:63-72 (comb).
//
// Log-weight by the saddle-point form (C. Loader, "Fast and accurate computation of binomial probabilities",
// 2000):  ln b(x;N,p) = d(N) - d(x) - d(N-x) - D(x;Np) - D(N-x;Nq) + N(p+q-1) - 1/2 ln(2 pi x (N-x)/N)
// with d(n) the Stirling error and D(x;m) = x ln(x/m) + m - x the deviance. Everything the two terms of a
// pair share is computed once, and the expensive scalar operations are batched:
//   * y = (x(N-x))^(-1/2) by MUFU.RSQ64H + two coupled iterations gives the prefactor (2 pi x(N-x)/N)^(-1/2)
//     AND 1/x, 1/(N-x) for the two Stirling series (y*y*(N-x), y*y*x): no logarithm, no division;
//   * the four deviances need 1/(x+Np), 1/(N-x+Nq), 1/(N-x+Np), 1/(x+Nq): one reciprocal of the product
//     and nine multiplications (Montgomery's trick) instead of four divisions;
//   * the deviance series in v = (x-m)/(x+m) stops as soon as a term is below 1e-17 of the sum (2-3 terms
//     where the weight is not negligible at large N); |v| >= 0.1 (far tails) uses x ln(x/m) with the table log.
//   * exponentials are exp_table (11 FP64 each).
// Np, Nq, ln u, ln d, ln p, ln q arrive as double-double from the host; u, d, p, q are the reference's own
// doubles (binom_embar.cpp:19-27).
#pragma once
#include "fastmath.cuh"

namespace pcf {

struct BinomArgs {
  double S0, E;
  int cp;
  double Nd;                 // N as double
  long long N;
  long long i0, i1;          // this GPU's pair range, i in [i0, i1), i < ceil(N/2)
  int add_mid;               // this GPU also adds the middle term N/2 (N even), as binom_embar.cpp:42-45
  double np_hi, np_lo, nq_hi, nq_lo;   // N*p, N*q as double-double
  double inv_np, inv_nq;               // 1/(N p), 1/(N q)
  double lnp_hi, lnp_lo, lnq_hi, lnq_lo;
  double lnu_hi, lnu_lo, lnd_hi, lnd_lo;
  double stirl_N;            // d(N)
  double corr;               // N*(p+q-1): q = fl(1-p) is not exactly 1-p, and the sum is defined on the
                             // reference's (p, q) doubles (binom_embar.cpp:24-27)
  double pref;               // sqrt(N / (2 pi))
  double ln_np, ln_nq;       // ln(N p), ln(N q): the screening pass (pair_dead)
  int screen;                // 1: pairs whose two weights provably underflow are settled by pair_dead()
  double sfe[16];            // d(0..15), host long-double values
};

#define PCF_INV_ODD {1.0 / 3, 1.0 / 5, 1.0 / 7, 1.0 / 9, 1.0 / 11, 1.0 / 13, 1.0 / 15, 1.0 / 17, 1.0 / 19}
#ifdef __CUDACC__
static __constant__ double c_inv_odd[9] = PCF_INV_ODD;
static __constant__ double c_sfe[16];  // d(0..15), uploaded by run_binom
#endif
static const double h_inv_odd[9] = PCF_INV_ODD;
#ifdef __CUDA_ARCH__
#define PCF_INVODD(j) c_inv_odd[j]
#define PCF_SFE(a, k) c_sfe[k]
#else
#define PCF_INVODD(j) h_inv_odd[j]
#define PCF_SFE(a, k) (a).sfe[k]
#endif

PCF_HD double payoff_hd(double St, double E, int cp) {
  double v = (double)cp * (St - E);
  return v > 0.0 ? v : 0.0;
}

// y = q^(-1/2) to ~1 ulp, q in [1, 2^62]
PCF_HD double rsqrt_pos(double q) {
#ifdef __CUDA_ARCH__
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(q));
#else
  double y0 = (double)(float)(1.0 / std::sqrt(q));
#endif
  double g = q * y0, h = 0.5 * y0;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-h, g, 0.5);
  h = fma(h, r, h);
  return h + h;
}

// 1/P to ~1 ulp, P in [1, 2^200]
PCF_HD double rcp_pos(double P) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(P));
#else
  double r = (double)(float)(1.0 / P);
#endif
  double e = fma(-P, r, 1.0);
  r = fma(r, e, r);
  e = fma(-P, r, 1.0);
  return fma(r, e, r);
}

// Stirling error d(n) = ln n! - [(n+1/2) ln n - n + 1/2 ln 2 pi]; inv_n = 1/n
PCF_HD double stirlerr(double n, double inv_n, const BinomArgs& a) {
  if (n < 16.0) return PCF_SFE(a, (int)n);
  const double r2 = inv_n * inv_n;
  // 1/12 - 1/360 r^2 + 1/1260 r^4 - 1/1680 r^6 + 1/1188 r^8 - 691/360360 r^10
  double s = -691.0 / 360360.0;
  s = fma(s, r2, 1.0 / 1188.0);
  s = fma(s, r2, -1.0 / 1680.0);
  s = fma(s, r2, 1.0 / 1260.0);
  s = fma(s, r2, -1.0 / 360.0);
  s = fma(s, r2, 1.0 / 12.0);
  return s * inv_n;
}

// deviance D(x; m), m = m_hi + m_lo, given inv_sum = 1/(x + m_hi) and inv_m = 1/m
PCF_HD double bd0(double x, double m_hi, double m_lo, double inv_m, double inv_sum, const TableView& tv,
                  const Hoisted& hc) {
  const double diff = (x - m_hi) - m_lo;
  const double sum = x + m_hi;
  if (fabs(diff) < 0.1 * sum) {
    const double v = diff * inv_sum, v2 = v * v;
    double s = diff * v;
    double ej = 2.0 * x * v;
#pragma unroll 1
    for (int j = 0; j < 9; ++j) {
      ej *= v2;
      const double t = ej * PCF_INVODD(j);
      s += t;
      if (!(fabs(t) > 1e-17 * s)) break;
    }
    return s;
  }
  const double ln_ratio = -0.5 * neg2log_unit(x * inv_m, tv, hc);
  return fma(x, ln_ratio, -diff);
}

// ln(S0 u^x d^(N-x)) - ln S0 = x ln u + (N-x) ln d, double-double products and sum
PCF_HD double log_growth(double x, double nx, const BinomArgs& a) {
  const double t1 = x * a.lnu_hi, e1 = fma(x, a.lnu_hi, -t1);
  const double t2 = nx * a.lnd_hi, e2 = fma(nx, a.lnd_hi, -t2);
  const double s = t1 + t2, bb = s - t1;
  const double err = (t1 - (s - bb)) + (t2 - bb);
  const double lo = err + e1 + e2 + fma(x, a.lnu_lo, nx * a.lnd_lo);
  return s + lo;
}

// weight * payoff; `lw` excludes the shared prefactor `rs` = (2 pi x (N-x)/N)^(-1/2)
PCF_HD double weighted_payoff(double lw, double rs, double x, double nx, const BinomArgs& a, const TableView& tv) {
  if (!(lw > -700.0)) return 0.0;  // the weight underflows: the term is 0, not 0*inf
  const double w = exp_table(lw, tv) * rs;
  const double S = a.S0 * exp_table(log_growth(x, nx, a), tv);
  return w * payoff_hd(S, a.E, a.cp);
}

// Screening pass. The saddle-point log-weight is lw = -D(x;Np) - D(N-x;Nq) + [d(N) - d(x) - d(N-x) + N(p+q-1)], the
// bracket is within 1/4 of 0 for every x >= 1, and weighted_payoff() defines a term with lw <= -700 as 0. Here both
// log-weights of the pair are evaluated from TWO logarithms,
//   lw1 ~ x (ln Np - ln x) + (N-x)(ln Nq - ln(N-x)),   lw2 ~ (N-x)(ln Np - ln(N-x)) + x (ln Nq - ln x)
// (the m - x parts of the four deviances cancel up to N(p+q-1)). Absolute error <= N * 22 * 2^-52 * 4 < 1e-4 for
// N < 2^31, so "both below -712" proves both terms are 0 by the rule above and pair_terms() -- which has to be accurate
// to 1e-16 where the weight is NOT negligible and costs four times as much -- would return exactly 0.0.
// At N = 1e8 all but ~4e5 of the 5e7 pairs end here.
PCF_HD bool pair_dead(double x, double nx, const BinomArgs& a, const TableView& tv, const Hoisted& hc) {
  const double lx = -0.5 * neg2log_unit(x, tv, hc), lnx = -0.5 * neg2log_unit(nx, tv, hc);
  const double lw1 = fma(x, a.ln_np - lx, nx * (a.ln_nq - lnx));
  const double lw2 = fma(nx, a.ln_np - lnx, x * (a.ln_nq - lx));
  return lw1 < -712.0 && lw2 < -712.0;
}

// terms i and N-i, 1 <= i <= N-i
PCF_HD double pair_terms(long long i, const BinomArgs& a, const TableView& tv, const Hoisted& hc) {
  const double x = (double)i, nx = (double)(a.N - i);
  const double y = rsqrt_pos(x * nx);
  const double inv_q = y * y;
  const double sx = stirlerr(x, nx * inv_q, a), snx = stirlerr(nx, x * inv_q, a);
  const double rs = y * a.pref;
  // reciprocals of the four (x + m) sums from one reciprocal
  const double s1 = x + a.np_hi, s2 = nx + a.nq_hi, s3 = nx + a.np_hi, s4 = x + a.nq_hi;
  const double p12 = s1 * s2, p34 = s3 * s4;
  const double R = rcp_pos(p12 * p34);
  const double r12 = R * p34, r34 = R * p12;
  const double D1 = bd0(x, a.np_hi, a.np_lo, a.inv_np, r12 * s2, tv, hc);
  const double D2 = bd0(nx, a.nq_hi, a.nq_lo, a.inv_nq, r12 * s1, tv, hc);
  const double D3 = bd0(nx, a.np_hi, a.np_lo, a.inv_np, r34 * s4, tv, hc);
  const double D4 = bd0(x, a.nq_hi, a.nq_lo, a.inv_nq, r34 * s3, tv, hc);
  const double common = (a.stirl_N + a.corr) - sx - snx;
  const double lw1 = common - D1 - D2;  // x ups
  const double lw2 = common - D3 - D4;  // N-x ups
  return weighted_payoff(lw1, rs, x, nx, a, tv) + weighted_payoff(lw2, rs, nx, x, a, tv);
}

// terms 0 and N: q^N and p^N (no Stirling form at the ends)
PCF_HD double end_terms(const BinomArgs& a, const TableView& tv) {
  const double lw1 = fma(a.Nd, a.lnq_hi, a.Nd * a.lnq_lo);
  const double lw2 = fma(a.Nd, a.lnp_hi, a.Nd * a.lnp_lo);
  return weighted_payoff(lw1, 1.0, 0.0, a.Nd, a, tv) + weighted_payoff(lw2, 1.0, a.Nd, 0.0, a, tv);
}

// ---- host side: lattice parameters and argument block --------------------------------------------------------
inline long double stirlerr_host(long double n) {
  if (n == 0) return 0.0L;
  const long double half_ln_2pi = 0.918938533204672741780329736405617639L;
  return lgammal(n + 1.0L) - ((n + 0.5L) * logl(n) - n + half_ln_2pi);
}

inline void split_ld(long double v, double& hi, double& lo) {
  hi = (double)v;
  lo = (double)(v - (long double)hi);
}

// Reference lattice parameters, binom_embar.cpp:19-27, evaluated with the same libm calls in the same order
// (the sqrt(beta^2-1) cancellation makes any algebraic rewrite visible).
inline void binom_lattice(double r, double sigma, double T, long long N, double& u, double& d, double& p, double& q) {
  double dt = (double)T / (double)N;
  double beta = 0.5 * (exp(-r * dt) + exp((r + sigma * sigma) * dt));
  u = beta + sqrt(beta * beta - 1);
  d = beta - sqrt(beta * beta - 1);
  double R = exp(r * dt);
  p = (R - d) / (u - d);
  q = 1 - p;
}

inline void fill_binom_args(double S0, double E, double r, double sigma, double T, long long N, int cp, BinomArgs& a) {
  double u, d, pp, q;
  binom_lattice(r, sigma, T, N, u, d, pp, q);
  a.S0 = S0; a.E = E; a.cp = cp; a.N = N; a.Nd = (double)N;
  a.i0 = 0; a.i1 = 0; a.add_mid = 0;
  a.np_hi = a.Nd * pp; a.np_lo = fma(a.Nd, pp, -a.np_hi);  // exact double-double products
  a.nq_hi = a.Nd * q;  a.nq_lo = fma(a.Nd, q, -a.nq_hi);
  a.corr = (double)((long double)N * (((long double)pp + (long double)q) - 1.0L));
  a.inv_np = (double)(1.0L / ((long double)N * (long double)pp));
  a.inv_nq = (double)(1.0L / ((long double)N * (long double)q));
  a.pref = (double)sqrtl((long double)N / 6.283185307179586476925286766559005768L);
  a.ln_np = (double)logl((long double)N * (long double)pp);
  a.ln_nq = (double)logl((long double)N * (long double)q);
  a.screen = 1;
  split_ld(logl((long double)pp), a.lnp_hi, a.lnp_lo);
  split_ld(logl((long double)q), a.lnq_hi, a.lnq_lo);
  split_ld(logl((long double)u), a.lnu_hi, a.lnu_lo);
  split_ld(logl((long double)d), a.lnd_hi, a.lnd_lo);
  if (N < 64) {
    a.stirl_N = (double)stirlerr_host((long double)N);
  } else {
    long double n = (long double)N, r2 = 1.0L / (n * n);
    a.stirl_N = (double)((1.0L / 12 - (1.0L / 360 - (1.0L / 1260 - (1.0L / 1680 - (1.0L / 1188) * r2) * r2) * r2) * r2) / n);
  }
  for (int k = 0; k < 16; ++k) a.sfe[k] = (double)stirlerr_host((long double)k);
}

}  // namespace pcf
