// TEST INFRASTRUCTURE. Function-level access to the UNMODIFIED reference pricing functions:
// the reference translation unit is #included where it lies (path given by -DREF_SRC), its
// main() renamed away, and the pricing function is called with argv parsed exactly as the
// reference's own main does (getArg/getArgD, include/common.h:25-38). Prints "%.17g\n".
// Built by oracle/Makefile with -DREF_mc_eur / -DREF_mc_eur_multi (against oracle/shim) / -DREF_mc_asia / -DREF_mc_amer / -DREF_binom_embar /
// -DREF_binom_vanilla_eur / -DREF_binom_vanilla_amer. The RNG seed is pinned by wrap_time.c (PCF_FIXED_TIME).
#define main ref_main_unused
#include REF_SRC
#undef main
#include <cstdio>

int main(int argc, char *argv[]) {
  std::string pf = argv[1];
  double cp = (pf == "call") ? 1 : -1;
  double S0 = getArgD(argv, 2), E = getArgD(argv, 3), r = getArgD(argv, 4);
  double sigma = getArgD(argv, 5), T = getArgD(argv, 6);
  int N = getArg(argv, 7);
  double res;
#if defined(REF_mc_eur_multi)
  res = mc_eur(S0, E, r, sigma, T, N, cp, getArg(argv, 8), getArgD(argv, 9));  // src/mc_eur_multi.cpp:56
#elif defined(REF_mc_eur)
  res = mc_eur(S0, E, r, sigma, T, N, cp);
#elif defined(REF_mc_asia)
  res = mc_asia(S0, E, r, sigma, T, N, getArg(argv, 8), cp);
#elif defined(REF_mc_amer)
  res = mc_amer(S0, E, r, sigma, T, N, getArg(argv, 8), cp);
#elif defined(REF_binom_embar) || defined(REF_binom_vanilla_eur) || defined(REF_binom_vanilla_amer)
  res = binom(S0, E, r, sigma, T, N, cp);
#endif
  std::printf("%.17g\n", res);
  return 0;
}
