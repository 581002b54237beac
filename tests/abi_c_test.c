/* Plain-C consumer of include/pcf.h (C99, -pedantic): the header is the reference-facing contract, so it must be valid C
 * and every struct must have the size the ctypes mirror assumes. Prints sizes, exercises the host-only entry points and
 * checks that a compute call without pcf_init() fails with PCF_ENOINIT (no CPU fallback). */
#include <stdio.h>
#include <string.h>
#include "pcf.h"

int main(void) {
  double L[16 * 16];
  double cov[4] = {1.0, 0.5, 0.5, 1.0}, A[4];
  int eig = -1, st;
  pcf_params p;
  pcf_result r;
  pcf_basket b;
  memset(&p, 0, sizeof p);
  memset(&r, 0, sizeof r);
  memset(&b, 0, sizeof b);
  printf("sizeof %u %u %u abi %d\n", (unsigned)sizeof(pcf_params), (unsigned)sizeof(pcf_result), (unsigned)sizeof(pcf_basket),
         PCF_ABI_VERSION);
  st = pcf_chol_equicorr(16, 0.5, L);
  printf("chol %d %.17g\n", st, L[16 * 15 + 15]);
  st = pcf_normal_transform(2, cov, A, &eig);
  printf("transform %d %d %.17g\n", st, eig, A[3]);
  p.S0 = 100; p.E = 100; p.r = 0.05; p.sigma = 0.2; p.T = 1; p.cp = 1; p.N = 1000; p.M = 10; p.assets = 1;
  st = pcf_mc_asia(&p, &r);
  printf("noinit %d %s\n", st, pcf_strerror(st));
  p.cp = 0;
  st = pcf_mc_eur(&p, &r);
  printf("badpayoff %d %s\n", st, pcf_strerror(st));
  return 0;
}
