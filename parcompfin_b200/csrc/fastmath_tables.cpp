// fastmath_tables.cpp -- host-side construction (long double) of the lookup tables of fastmath.cuh.
#include <cmath>
#include "fastmath.cuh"

namespace pcf {

void build_math_tables(MathTables& t) {
  for (int j = 0; j < kLnEntries; ++j) {
    // bin j = re-based high words [j*2^13, (j+1)*2^13) above 0x3FE6A09E (see neg2log_unit)
    const double lo = make_double(0x3FE6A09Eu + (uint32_t)j * 8192u, 0u);
    const double hi = make_double(0x3FE6A09Eu + (uint32_t)(j + 1) * 8192u, 0u);
    const long double centre = ((long double)lo + (long double)hi) / 2.0L;
    const double rc = (double)(1.0L / centre);
    t.ln_tab[j].x = rc;
    t.ln_tab[j].y = (double)(2.0L * logl((long double)rc) + 0x1p-56L);
  }
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int j = 0; j < kScEntries; ++j) {
    const long double a = two_pi * ((long double)j + 0.5L) / (long double)kScEntries;
    t.sc_tab[j].x = (double)cosl(a);
    t.sc_tab[j].y = (double)sinl(a);
  }
  for (int j = 0; j < kExpEntries; ++j) t.exp_tab[j] = (double)exp2l((long double)j / (long double)kExpEntries);
}

}  // namespace pcf
