// Host evaluation of parcompfin_b200/csrc/binom_math.cuh (the routines are __host__ __device__): sums the
// term pairs exactly like binom_terms_kernel does and prints "%.17g" prices for the cases on the command line:
//   binom_host_test <call|put> S0 E r sigma T N ...   (7 arguments per case)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../parcompfin_b200/csrc/binom_math.cuh"
namespace pcf { void build_math_tables(MathTables& t); }
using namespace pcf;

int main(int argc, char** argv) {
  static MathTables T;
  build_math_tables(T);
  TableView tv{T.ln_tab, T.sc_tab, T.exp_tab, 1, 1};
  Hoisted hc;
  hc.load();
  for (int k = 1; k + 6 < argc; k += 7) {
    int cp = std::strcmp(argv[k], "call") == 0 ? 1 : -1;
    double S0 = atof(argv[k + 1]), E = atof(argv[k + 2]), r = atof(argv[k + 3]), sigma = atof(argv[k + 4]), Tm = atof(argv[k + 5]);
    long long N = atoll(argv[k + 6]);
    BinomArgs a;
    fill_binom_args(S0, E, r, sigma, Tm, N, cp, a);
    long long until = (N % 2 != 0) ? (N + 1) / 2 : N / 2;
    long double sum = end_terms(a, tv);
    // the screening pass (pair_dead) may only settle pairs that pair_terms evaluates to exactly 0.0
    long long dead = 0, wrong = 0;
    for (long long i = 1; i < until; ++i) {
      const double t = pair_terms(i, a, tv, hc);
      if (pair_dead((double)i, (double)(N - i), a, tv, hc)) {
        ++dead;
        if (t != 0.0) ++wrong;
      }
      sum += t;
    }
    if (wrong) { std::printf("SCREEN VIOLATION %lld of %lld\n", wrong, dead); return 1; }
    std::fprintf(stderr, "N=%lld screened %lld of %lld pairs\n", N, dead, until - 1);
    if (N % 2 == 0) sum += 0.5 * pair_terms(N / 2, a, tv, hc);
    std::printf("%.17g\n", (double)(expl(-(long double)r * Tm) * sum));
  }
  return 0;
}
