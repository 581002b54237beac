// fastmath.cuh -- FP64 elementary functions sized for the FP64-pipe budget of the Monte Carlo kernels.
//
// The CUDA math library's log / sincospi / sqrt / exp cost 30 / 21 / 9 / 17 FP64-pipe instructions and
// pull ~50 UMOV constant materialisations per loop iteration (profiles/r1_*). The routines
// here reach the same <= 1-2 ulp accuracy with small shared-memory tables (bank-replicated so that a
// warp's 32 random lookups are conflict-free) and short polynomials whose coefficients are
// constant-bank operands:
//   neg2log_unit(u)     -2 ln u, u in (0,1]            10 FP64   (128-entry table of 1/c, 2 ln(1/c))
//   sqrt_pos(t)         sqrt t, t > 0                    7 FP64   (MUFU.RSQ64H seed + coupled iteration)
//   sincos_sector(...)  cos/sin of 2 pi u                14 FP64   (64 sector table + degree 7/8 polys)
//   exp_small(x)        e^x, |x| <= 0.11                  9 FP64   (Taylor degree 9, no reduction)
//   exp_table(x)        e^x, any finite x                11 FP64   (32-entry 2^(j/32) table, degree 6)
// All are __host__ __device__ so that tests/fastmath_host_test.cpp can sweep them against long double
// libm on the CPU; on the host the tables are indexed without replication.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>

#ifdef __CUDACC__
#define PCF_HD __host__ __device__ __forceinline__
#else
#define PCF_HD inline
#endif

namespace pcf {

constexpr int kLnEntries = 128;
constexpr int kScEntries = 64;
constexpr int kExpEntries = 32;
constexpr int kRep16 = 8;   // replicas of a 16-byte-entry table: entry (idx, lane & 7) -> conflict-free LDS.128
constexpr int kRep8 = 16;   // replicas of an 8-byte-entry table:  entry (idx, lane & 15) -> conflict-free LDS.64

struct Pair {
  double x, y;
};

// Unreplicated tables, built on the host in long double (fastmath_tables.cpp), uploaded once per context.
struct MathTables {
  Pair ln_tab[kLnEntries];   // x = rc_j = fl(1/center_j), y = 2 ln(rc_j) + 2^-56
  Pair sc_tab[kScEntries];   // x = cos(2 pi (j + 1/2)/64), y = sin(...)
  double exp_tab[kExpEntries];  // 2^(j/32)
};

constexpr size_t kTableSmemBytes = (size_t)kLnEntries * kRep16 * 16 + (size_t)kScEntries * kRep16 * 16 +
                                   (size_t)kExpEntries * kRep8 * 8;  // 16 KB + 8 KB + 4 KB

// Per-thread view of the (replicated) tables.
struct TableView {
  const Pair* ln_tab;     // pre-offset by the lane's replica
  const Pair* sc_tab;
  const double* exp_tab;
  int stride16, stride8;  // kRep16 / kRep8 on the device, 1 on the host
};

// Polynomial coefficients and constants. On the device they live in the constant bank, so every DFMA takes
// its coefficient as a c[bank][offset] operand instead of a pair of UMOV-materialised immediates.
enum CoefIndex {
  kLn0, kLn1, kLn2, kLn3, kLn4, kLnEMagic, kLnNeg2Ln2,          // neg2log_unit
  kScA, kScB, kS7, kS5, kS3, kC8, kC6, kC4, kC2,                // sincos_2pi_bits
  kE9, kE8, kE7, kE6, kE5, kE4, kE3, kE2,                       // exp_small (1/9! .. 1/2!)
  kE10,                                                         // exp_small_pm
  kXMagic, kX32Ln2, kXLn2Hi, kXLn2Lo,                           // exp_table
  kHalf, kOne, kNegOne, kNegTwo,
  kCoefCount
};
#define PCF_COEF_VALUES                                                                                      \
  {1.0 / 3.0, -2.0 / 5.0, 0.5, -2.0 / 3.0, 1.0, 4503599627371519.0, -1.3862943611198906,                     \
   0.09817477042468103, -0.14726215563702155, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0, 1.0 / 40320.0,          \
   -1.0 / 720.0, 1.0 / 24.0, -0.5,                                                                           \
   1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5,        \
   1.0 / 3628800.0,                                                                                          \
   6755399441055744.0, 46.16624130844683, -0.02166084939249829, -7.247021293269686e-19,                      \
   0.5, 1.0, -1.0, -2.0}
#ifdef __CUDACC__
static __constant__ double c_coef[kCoefCount] = PCF_COEF_VALUES;
#endif
static const double h_coef[kCoefCount] = PCF_COEF_VALUES;
#ifdef __CUDA_ARCH__
#define K(i) c_coef[i]
#else
#define K(i) h_coef[i]
#endif

// A DFMA takes at most one constant-bank operand, so the first Horner step of every polynomial
// (two coefficients) needs one of them in a vector register. Loading these once per thread, before the hot
// loop, keeps LDC (and its scoreboard wait) out of the loop body.
struct Hoisted {
  double ln1, s5, c6, e8, e7;
  PCF_HD void load() {
    ln1 = K(kLn1);
    s5 = K(kS5);
    c6 = K(kC6);
    e8 = K(kE8);
    e7 = K(kE7);
  }
};

PCF_HD double make_double(uint32_t hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double((int)hi, (int)lo);
#else
  uint64_t b = ((uint64_t)hi << 32) | lo;
  double d;
  std::memcpy(&d, &b, 8);
  return d;
#endif
}
PCF_HD uint32_t hi_word(double d) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__double2hiint(d);
#else
  uint64_t b;
  std::memcpy(&b, &d, 8);
  return (uint32_t)(b >> 32);
#endif
}
PCF_HD uint32_t lo_word(double d) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__double2loint(d);
#else
  uint64_t b;
  std::memcpy(&b, &d, 8);
  return (uint32_t)b;
#endif
}

// ---- -2 ln u for u in [2^-60, 1] ---------------------------------------------------------------------
// u = 2^e m with m in [sqrt(1/2), sqrt(2)) (so that e = 0 around u = 1: no cancellation against e ln 2);
// bin j = bits 19..13 of the re-based high word, r = m * rc_j - 1 (one FMA, |r| <= 2^-8),
//   -2 ln u = e (-2 ln 2) + 2 ln rc_j - 2 log1p(r),  -2 log1p(r) = r(-2 + r(1 + r(-2/3 + r(1/2 + r(-2/5 + r/3)))))
// The table's 2^-56 bias keeps the result strictly positive at u = 1 (its absolute error is ~1e-16 anyway).
PCF_HD double neg2log_unit(double u, const TableView& tv, const Hoisted& hc) {
  const uint32_t hx = hi_word(u) + 0x00095F62u;                 // 0x3FF00000 - 0x3FE6A09E
  const uint32_t idx = (hx >> 13) & 0x7Fu;
  const Pair e = tv.ln_tab[idx * tv.stride16];
  const double m = make_double((hx & 0x000FFFFFu) + 0x3FE6A09Eu, lo_word(u));
  const double ed = make_double(0x43300000u, hx >> 20) - K(kLnEMagic);  // (2^52 + biased) - (2^52 + 1023)
  const double r = fma(m, e.x, K(kNegOne));
  double q = fma(r, K(kLn0), hc.ln1);    // 1/3, -2/5
  q = fma(q, r, K(kLn2));                // 1/2
  q = fma(q, r, K(kLn3));                // -2/3
  q = fma(q, r, K(kLn4));                // 1
  q = fma(q, r, K(kNegTwo));
  const double base = fma(ed, K(kLnNeg2Ln2), e.y);  // -2 ln 2
  return fma(q, r, base);
}

// ---- sqrt(t), t in [2^-200, 2^200] -------------------------------------------------------------------
PCF_HD double sqrt_pos(double t) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));      // MUFU.RSQ64H: ~2^-22 relative
#else
  double y = (double)(float)(1.0 / std::sqrt(t));               // same accuracy class as the hardware seed
#endif
  double g = t * y;                // ~ sqrt(t)
  double h = K(kHalf) * y;         // ~ 1 / (2 sqrt(t))
  double r = fma(-h, g, K(kHalf));
  g = fma(g, r, g);
  h = fma(h, r, h);
  double d = fma(-g, g, t);
  return fma(d, h, g);
}

// ---- cos/sin(2 pi u2), u2 = ((X2 >> 6) + 1/2) 2^-58, X2 = x3:x2 given as the two Philox words ----------
// sector = top 6 bits of X2 (table of cos/sin at the sector centre); the next 52 bits f fill a mantissa,
// w = 1 + f 2^-52, and the offset from the centre is x = (pi/32)(w - 3/2 + 2^-53), |x| < pi/64:
// sin to x^7, cos to x^8, then one rotation.
PCF_HD void sincos_2pi_bits(uint32_t x2, uint32_t x3, const TableView& tv, const Hoisted& hc, double& c_out,
                            double& s_out) {
  const uint32_t sector = x3 >> 26;
  const Pair cs = tv.sc_tab[sector * tv.stride16];
  const double w = make_double(0x3FF00000u | ((x3 >> 6) & 0x000FFFFFu), (x3 << 26) | (x2 >> 6));
  const double x = fma(w, K(kScA), K(kScB));  // pi/32, -(pi/32)(3/2 - 2^-53)
  const double x2d = x * x;
  double ps = fma(x2d, K(kS7), hc.s5);
  ps = fma(ps, x2d, K(kS3));
  const double x3d = x * x2d;
  const double s = fma(x3d, ps, x);
  double pc = fma(x2d, K(kC8), hc.c6);
  pc = fma(pc, x2d, K(kC4));
  pc = fma(pc, x2d, K(kC2));
  const double c = fma(pc, x2d, K(kOne));
  c_out = fma(cs.x, c, -cs.y * s);
  s_out = fma(cs.y, c, cs.x * s);
}

// ---- e^x for |x| <= 0.11 (per-step GBM increments): Taylor degree 9, truncation < 2.6e-17 relative ------
PCF_HD double exp_small(double x, const Hoisted& hc) {
  double p = fma(x, K(kE9), hc.e8);
  p = fma(p, x, K(kE7));
  p = fma(p, x, K(kE6));
  p = fma(p, x, K(kE5));
  p = fma(p, x, K(kE4));
  p = fma(p, x, K(kE3));
  p = fma(p, x, K(kE2));
  p = fma(p, x, K(kOne));
  return fma(p, x, K(kOne));
}

// ---- e^x, |x| < 700: x = (32 k + j) ln2/32 + r, |r| <= ln2/64; e^x = 2^k 2^(j/32) (1 + r q(r)) ---------
PCF_HD double exp_table(double x, const TableView& tv) {
  const double magic = K(kXMagic);                                // 1.5 * 2^52: integer lands in the low word
  const double t = fma(x, K(kX32Ln2), magic);                     // 32 / ln 2
  const double kf = t - magic;
  double r = fma(kf, K(kXLn2Hi), x);                              // ln2/32 = hi + lo
  r = fma(kf, K(kXLn2Lo), r);
  const uint32_t n = lo_word(t);
  const double T = tv.exp_tab[(n & 31u) * tv.stride8];
  double q = fma(r, K(kE6), K(kE5));
  q = fma(q, r, K(kE4));
  q = fma(q, r, K(kE3));
  q = fma(q, r, K(kE2));
  q = fma(q, r, K(kOne));
  const double rq = r * q;
  const double v = fma(T, rq, T);
  // scale by 2^k, k = n >> 5 (arithmetic): add k to the exponent field
  const int k = (int)n >> 5;
  return make_double(hi_word(v) + ((uint32_t)k << 20), lo_word(v));
}

// e^(a+x) and e^(a-x) for |x| <= 0.11 from one even/odd split (antithetic pairs): 13 FP64 for both.
PCF_HD void exp_small_pm(double x, const Hoisted& hc, double& ep, double& em) {
  const double x2 = x * x;
  double ce = fma(x2, K(kE10), hc.e8);
  ce = fma(ce, x2, K(kE6));
  ce = fma(ce, x2, K(kE4));
  ce = fma(ce, x2, K(kE2));
  ce = fma(ce, x2, K(kOne));
  double so = fma(x2, K(kE9), hc.e7);
  so = fma(so, x2, K(kE5));
  so = fma(so, x2, K(kE3));
  so = fma(so, x2, K(kOne));
  so *= x;
  ep = ce + so;
  em = ce - so;
}

#undef K
}  // namespace pcf
