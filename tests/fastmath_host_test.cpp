// Host sweep of parcompfin_b200/csrc/fastmath.cuh against long double libm (the routines are
// __host__ __device__). Built and run by tests/test_fastmath_host.py; prints max errors.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../parcompfin_b200/csrc/fastmath.cuh"
namespace pcf { void build_math_tables(MathTables& t); }
using namespace pcf;

static double ulp_of(double x) { return std::nextafter(std::fabs(x), INFINITY) - std::fabs(x); }

int main() {
  static MathTables T;
  build_math_tables(T);
  TableView tv{T.ln_tab, T.sc_tab, T.exp_tab, 1, 1};
  Hoisted hc;
  hc.load();
  std::mt19937_64 g(12345);
  const int n = 2000000;
  // -2 ln u on the stream's own grid u = 1 - a 2^-52, incl. the ends
  double e_abs = 0, e_min = 1e300;
  for (int i = 0; i < n + 64; ++i) {
    uint64_t a = g() >> 12;
    if (i < 32) a = (uint64_t)i;                                // u = 1, 1 - 2^-52, ...
    else if (i < 64) a = (1ull << 52) - 1 - (uint64_t)(i - 32); // u = 2^-52, ...
    else if (i & 1) a >>= (g() % 50);                           // many magnitudes of 1 - u
    double u = 1.0 - (double)a * 0x1p-52;
    double got = neg2log_unit(u, tv, hc);
    long double want = -2.0L * logl((long double)u);
    double err = (double)fabsl((long double)got - want);
    if (err > e_abs) e_abs = err;
    if (got < e_min) e_min = got;
  }
  std::printf("neg2log_abs_err %.3e\nneg2log_min %.3e\n", e_abs, e_min);
  // sqrt
  double e_sqrt = 0;
  for (int i = 0; i < n; ++i) {
    double t = std::ldexp((double)(g() >> 11) * 0x1p-53 + 0.5, (int)(g() % 70) - 60);
    double got = sqrt_pos(t);
    long double want = sqrtl((long double)t);
    double err = (double)fabsl((long double)got - want) / ulp_of((double)want);
    if (err > e_sqrt) e_sqrt = err;
  }
  std::printf("sqrt_ulp %.3f\n", e_sqrt);
  // sincos
  double e_sc = 0, e_norm = 0;
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int i = 0; i < n; ++i) {
    uint64_t X2 = g();
    if (i < 64) X2 = ((uint64_t)i << 58);
    else if (i < 128) X2 = ((uint64_t)(i - 64) << 58) | ((1ull << 58) - 1);
    double c, s;
    sincos_2pi_bits((uint32_t)X2, (uint32_t)(X2 >> 32), tv, hc, c, s);
    long double u2 = ((long double)(X2 >> 6) + 0.5L) * 0x1p-58L;
    double ec = (double)fabsl((long double)c - cosl(two_pi * u2)), es = (double)fabsl((long double)s - sinl(two_pi * u2));
    if (ec > e_sc) e_sc = ec;
    if (es > e_sc) e_sc = es;
    double nn = std::fabs(c * c + s * s - 1.0);
    if (nn > e_norm) e_norm = nn;
  }
  std::printf("sincos_abs_err %.3e\nsincos_norm_err %.3e\n", e_sc, e_norm);
  // exp_small, exp_table, exp_small_pm
  double e_es = 0, e_et = 0, e_pm = 0;
  for (int i = 0; i < n; ++i) {
    double x = ((double)(g() >> 11) * 0x1p-53 - 0.5) * 0.22;
    double got = exp_small(x, hc);
    long double want = expl((long double)x);
    double err = (double)fabsl((long double)got - want) / ulp_of((double)want);
    if (err > e_es) e_es = err;
    double ep, em;
    exp_small_pm(x, hc, ep, em);
    err = (double)fabsl((long double)ep - want) / ulp_of((double)want);
    if (err > e_pm) e_pm = err;
    err = (double)fabsl((long double)em - expl(-(long double)x)) / ulp_of((double)expl(-(long double)x));
    if (err > e_pm) e_pm = err;
    double y = ((double)(g() >> 11) * 0x1p-53 - 0.5) * ((i & 1) ? 1400.0 : 6.0);
    got = exp_table(y, tv);
    want = expl((long double)y);
    err = (double)fabsl((long double)got - want) / ulp_of((double)want);
    if (err > e_et) e_et = err;
  }
  std::printf("exp_small_ulp %.3f\nexp_small_pm_ulp %.3f\nexp_table_ulp %.3f\n", e_es, e_pm, e_et);
  return 0;
}
