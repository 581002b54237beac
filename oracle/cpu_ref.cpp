// oracle/cpu_ref.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (plain FP64 C++, no FMA contraction) of the Monte Carlo / binomial hot path of
// moledoc/parcompfin. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load this library (liboracle.so); the product (libpcf.so) never
// does and has no CPU fallback.
//
// Parity status of this restatement:
//   * mc_eur, mc_asia, mc_amer, binom_embar: PINNED. tests/test_oracle.py checks every function
//     against the unmodified reference compiled from /root/reference (oracle/_ref/*_fn, seed pinned
//     through --wrap=time) -- see tests/golden/reference_vectors.json + tests/golden/make_golden.py
//     -- and against the reference's own published results/results_binom_embar.csv rows.
//   * mc_basket (reference src/mc_eur_multi.cpp + include/mvn.h): PINNED against the unmodified reference
//     translation unit compiled with the stand-in Eigen / Boost.Random headers of oracle/shim/ (the image has
//     neither library and the reference pins no version of them): oracle/_ref/mc_eur_multi_fn. That pins
//     everything the reference's own code decides -- covariance build, Cholesky-or-eigen branch, draw order
//     Z[n*d + a], the product normTransform * Z, the missing sqrt(T), weights, payoff, discount.
//     The factorisations underneath (LLT, SelfAdjointEigenSolver) are the stand-in's, used here through the
//     same API calls mvn.h makes; Eigen's own kernels could differ from them in the last bits.
//
// Every function takes the normal variates as an input array ("replay stream", reference draw
// order) so that the same stream can be fed to the CUDA kernels.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>
#include <omp.h>
#include "shim/Eigen/Dense"  // the same stand-in LLT / SelfAdjointEigenSolver the compiled reference links (oracle/shim)

#define ORACLE_API extern "C" __attribute__((visibility("default")))

namespace {

// reference include/common.h:56-60 -- cp arrives as a double, is truncated to int at the call.
inline double payoff(double St, double E, int cp) { return std::max(cp * (St - E), 0.0); }

}  // namespace

// ---------------------------------------------------------------------------------------------
// Reference RNG: std::mt19937 seeded with `seed`, one std::normal_distribution<>{0,sd} object that
// lives for the whole run (libstdc++ polar method, caches the second variate).
// reference src/mc_eur.cpp:16-20, src/mc_asia.cpp:20-24, include/common.h:188-192.
ORACLE_API void oracle_normals_mt19937(uint64_t seed, double sd, long long n, double* out) {
  std::mt19937 gen;
  gen.seed((std::mt19937::result_type)seed);
  std::normal_distribution<> norm{0, sd};
  for (long long i = 0; i < n; ++i) out[i] = norm(gen);
}

// ---------------------------------------------------------------------------------------------
// a1: reference src/mc_eur.cpp:5-27.  w[n] ~ N(0, T) in draw order.
ORACLE_API double oracle_mc_eur(double S0, double E, double r, double sigma, double T, long long N,
                                int cp, const double* w, double* sum_out, double* sumsq_out) {
  double acc = 0, acc2 = 0;
  const double drift = (r - sigma * sigma / 2) * T;  // pow(sigma,2)/2 -- mc_eur.cpp:24
  for (long long n = 0; n < N; ++n) {
    double v = payoff(S0 * std::exp(drift + sigma * w[n]), E, cp);
    acc += v;
    acc2 += v * v;
  }
  if (sum_out) *sum_out = acc;
  if (sumsq_out) *sumsq_out = acc2;
  return (std::exp(-r * T) * acc) / (double)N;  // mc_eur.cpp:26
}

// ---------------------------------------------------------------------------------------------
// a3: reference src/mc_asia.cpp:5-40.  dB[n*M + m] ~ N(0, dt).
ORACLE_API double oracle_mc_asia(double S0, double E, double r, double sigma, double T, long long N,
                                 int M, int cp, const double* dB, double* sum_out,
                                 double* sumsq_out) {
  const double dt = (double)T / (double)M;  // mc_asia.cpp:17
  double acc = 0, acc2 = 0;
  for (long long n = 0; n < N; ++n) {
    double St = S0, I = 0;
    const double* z = dB + n * (long long)M;
    for (int m = 0; m < M; ++m) {
      double dBi = z[m];
      I += St * (1 + r * dt / 2 + sigma * dBi / 2);               // mc_asia.cpp:33 (pre-update St)
      St *= std::exp((r - sigma * sigma / 2) * dt + sigma * dBi);  // mc_asia.cpp:34
    }
    double v = payoff(I / (double)M, E, cp);  // mc_asia.cpp:36
    acc += v;
    acc2 += v * v;
  }
  if (sum_out) *sum_out = acc;
  if (sumsq_out) *sumsq_out = acc2;
  return std::exp(-r * T) * acc / (double)N;  // mc_asia.cpp:39
}

// ---------------------------------------------------------------------------------------------
// a6: reference include/common.h:168-208 (pathsfinder). Time-major (M+1) x N, antithetic halves.
// w[p*M + (m-1)] ~ N(0, dt) for pair p < N/2.  Returns 0, or 1 if N is odd (common.h:180 throws).
ORACLE_API int oracle_pathsfinder(double S0, double r, double sigma, double T, long long N, int M,
                                  const double* w, double* paths /* (M+1)*N */) {
  if (N % 2 != 0) return 1;
  const double dt = T / M;
  const long long H = N / 2;
  for (long long p = 0; p < H; ++p) {
    paths[p] = S0;
    paths[p + H] = S0;
    for (int m = 1; m <= M; ++m) {
      double x = w[p * (long long)M + (m - 1)];
      paths[m * N + p] = paths[(m - 1) * N + p] * std::exp((r - 0.5 * sigma * sigma) * dt + sigma * x);
      paths[m * N + p + H] =
          paths[(m - 1) * N + p + H] * std::exp((r - 0.5 * sigma * sigma) * dt - sigma * x);
    }
  }
  return 0;
}

namespace {
// a8: reference include/common.h:98-141 (inverse + mat_vec_mul), same operation order.
// Returns false when the determinant is <= 0 (common.h:115-117 throws).
bool solve3(const double x[3][3], const double y[3], double coef[3]) {
  double det = 0;
  for (int i = 0; i < 3; ++i)
    det += (x[0][i] * (x[1][(i + 1) % 3] * x[2][(i + 2) % 3] - x[1][(i + 2) % 3] * x[2][(i + 1) % 3]));
  if (det <= 0) return false;
  double inv[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      inv[j][i] = ((x[(j + 1) % 3][(i + 1) % 3] * x[(j + 2) % 3][(i + 2) % 3]) -
                   (x[(j + 1) % 3][(i + 2) % 3] * x[(j + 2) % 3][(i + 1) % 3])) / det;
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int j = 0; j < 3; ++j) s += inv[i][j] * y[j];
    coef[i] = s;
  }
  return true;
}
}  // namespace

// a7: reference src/mc_amer.cpp:5-114 -- the reference's own early-exercise scheme, quirks kept:
//   * regressor x = S - E (:47-48), exercise test payoff(x,E) > Yhat on the SHIFTED value (:100)
//     and that value is booked as the cash flow (:103)                               [SURVEY F1]
//   * continuation value uses the TRUE payoff at paths[when][n] (:50)
//   * x == -1 is the "out of the money" sentinel (:32,:98): an ITM path with S-E == -1 is skipped
//   * <= 2 ITM paths: compare true payoff with cont, book the true payoff (:75-83); 0: skip (:73)
// status: 0 ok, 1 odd N, 2 determinant <= 0.
static double mc_amer_impl(double S0, double E, double r, double sigma, double T, long long N, int M, int cp,
                           const double* w, int* status, bool lsm);

ORACLE_API double oracle_mc_amer(double S0, double E, double r, double sigma, double T, long long N,
                                 int M, int cp, const double* w, int* status) {
  return mc_amer_impl(S0, E, r, sigma, T, N, M, cp, w, status, false);
}

// the same scheme with the textbook Longstaff-Schwartz decision (true payoff against the
// fitted continuation value, true payoff booked, no sentinel skip) -- what depr/mc_amer/v3/mc_amer.cpp:84
// computed before `x` was shifted. Used to check the product's PCF_FLAG_AMER_LSM mode.
ORACLE_API double oracle_mc_amer_lsm(double S0, double E, double r, double sigma, double T, long long N,
                                     int M, int cp, const double* w, int* status) {
  return mc_amer_impl(S0, E, r, sigma, T, N, M, cp, w, status, true);
}

static double mc_amer_impl(double S0, double E, double r, double sigma, double T, long long N, int M, int cp,
                           const double* w, int* status, bool lsm) {
  *status = 0;
  std::vector<double> paths((size_t)(M + 1) * (size_t)N);
  if (oracle_pathsfinder(S0, r, sigma, T, N, M, w, paths.data())) {
    *status = 1;
    return NAN;
  }
  const double dt = T / M;
  std::vector<double> when((size_t)N, (double)M), st((size_t)N);
  for (long long n = 0; n < N; ++n) st[n] = payoff(paths[(size_t)M * N + n], E, cp);

  std::vector<double> x((size_t)N), y((size_t)N);
  for (int m = M - 1; m > 0; --m) {
    std::fill(x.begin(), x.end(), -1.0);
    std::fill(y.begin(), y.end(), -1.0);
    double sx = 0, sx2 = 0, sx3 = 0, sx4 = 0, sy = 0, syx = 0, syx2 = 0, cnt = 0;
    double fpo[2], fy[2];
    long long fn[2];
    int nf = 0;
    for (long long n = 0; n < N; ++n) {
      double pv = payoff(paths[(size_t)m * N + n], E, cp);
      if (pv > 0) {
        ++cnt;
        double ex = paths[(size_t)m * N + n] - E;
        x[n] = ex;
        double cont = std::exp(-r * dt * (when[n] - m)) *
                      payoff(paths[(size_t)when[n] * N + n], E, cp);
        y[n] = cont;
        sx += ex;
        sx2 += ex * ex;
        sx3 += ex * ex * ex;
        sx4 += ex * ex * ex * ex;
        sy += cont;
        syx += cont * ex;
        syx2 += cont * ex * ex;
        if (cnt <= 2) {
          fpo[nf] = pv;
          fy[nf] = cont;
          fn[nf] = n;
          ++nf;
        }
      }
    }
    if (cnt == 0) continue;
    if (cnt <= 2) {
      for (int i = 0; i < nf; ++i)
        if (fpo[i] > fy[i]) {
          when[fn[i]] = m;
          st[fn[i]] = fpo[i];
        }
      continue;
    }
    const double A[3][3] = {{cnt, sx, sx2}, {sx, sx2, sx3}, {sx2, sx3, sx4}};
    const double b[3] = {sy, syx, syx2};
    double coef[3];
    if (!solve3(A, b, coef)) {
      *status = 2;
      return NAN;
    }
    for (long long i = 0; i < N; ++i) {
      const bool itm = lsm ? (payoff(paths[(size_t)m * N + i], E, cp) > 0) : (x[i] != -1);
      if (itm) {
        double yhat = coef[0] + coef[1] * x[i] + coef[2] * (x[i] * x[i]);  // pow(x,2) :99
        double pv = lsm ? payoff(paths[(size_t)m * N + i], E, cp)          // textbook rule
                        : payoff(x[i], E, cp);                             // :100 (shifted value)
        if (pv > yhat) {
          when[i] = m;
          st[i] = pv;
        }
      }
    }
  }
  double result = 0;
  for (long long n = 0; n < N; ++n)
    if (st[n] != 0) result += std::exp(-r * when[n] * dt) * st[n];  // :110
  return std::max(payoff(S0, E, cp), result / (double)N);           // :113
}

// ---------------------------------------------------------------------------------------------
// a4: reference include/mvn.h:53-70. Equicorrelation matrix, lower Cholesky factor L (row-major
// d x d, upper part zero). Returns 0, or 3 if the matrix is not positive definite (the reference
// falls back to an eigen-decomposition there, mvn.h:72-76; this restatement reports it instead).
ORACLE_API int oracle_chol_equicorr(int d, double rho, double* L) {
  std::vector<double> C((size_t)d * d);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) C[(size_t)i * d + j] = (i != j) ? rho : 1.0;
  std::memset(L, 0, sizeof(double) * (size_t)d * d);
  for (int i = 0; i < d; ++i) {
    for (int j = 0; j <= i; ++j) {
      double s = C[(size_t)i * d + j];
      for (int k = 0; k < j; ++k) s -= L[(size_t)i * d + k] * L[(size_t)j * d + k];
      if (i == j) {
        if (!(s > 0)) return 3;
        L[(size_t)i * d + i] = std::sqrt(s);
      } else {
        L[(size_t)i * d + j] = s / L[(size_t)j * d + j];
      }
    }
  }
  return 0;
}

// a5: reference src/mc_eur_multi.cpp:6-35 with Bt = L * Z (mvn.h:78-80), Z[n*d + a] iid N(0,1).
// NOTE no sqrt(T) on the Brownian term (mc_eur_multi.cpp:30)                       [SURVEY F9]
static double basket_core(double S0, double E, double r, double sigma, double T, long long N, int cp,
                          int d, const double* L, const double* Z, double* sum_out,
                          double* sumsq_out, bool threaded) {
  const double w_i = 1.0 / (double)d;
  double acc = 0, acc2 = 0;
#pragma omp parallel if (threaded)
  {
    std::vector<double> bt((size_t)d);
#pragma omp for schedule(dynamic, 1000) reduction(+ : acc, acc2) nowait
    for (long long n = 0; n < N; ++n) {
      const double* z = Z + n * (long long)d;
      for (int a = 0; a < d; ++a) {
        double s = 0;
        for (int k = 0; k <= a; ++k) s += L[(size_t)a * d + k] * z[k];
        bt[a] = s;
      }
      double basket = 0;
      for (int a = 0; a < d; ++a)
        basket += w_i * S0 * std::exp((r - sigma * sigma / 2) * T + sigma * bt[a]);
      double v = payoff(basket, E, cp);
      acc += v;
      acc2 += v * v;
    }
  }
  if (sum_out) *sum_out = acc;
  if (sumsq_out) *sumsq_out = acc2;
  return (std::exp(-r * T) * acc) / (double)N;
}

ORACLE_API double oracle_mc_basket_general(const double* S0, double E, double r, const double* sigma, double T,
                                           long long N, int cp, int d, const double* A, const double* w,
                                           const double* Z, double* sum_out, double* sumsq_out);

// a4: reference include/mvn.h:53-76 call for call: covar (1 on the diagonal, rho elsewhere), LLT, and when that
// reports a non-positive pivot the eigen-decomposition fallback eigenvectors * sqrt(eigenvalues) (NaN entries where an
// eigenvalue of the semi-definite matrix comes out as -1e-17: that is what the reference computes). A is row-major
// d x d; *used_eigen says which branch was taken.
ORACLE_API void oracle_mvn_transform(int d, double rho, double* A, int* used_eigen) {
  Eigen::MatrixXd covar(d, d);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) covar(i, j) = (i != j) ? rho : 1;
  Eigen::MatrixXd normTransform(d, d);
  Eigen::LLT<Eigen::MatrixXd> cholSolver(covar);
  if (cholSolver.info() == Eigen::Success) {
    normTransform = cholSolver.matrixL();
    *used_eigen = 0;
  } else {
    Eigen::SelfAdjointEigenSolver<Eigen::MatrixXd> eigenSolver(covar);
    normTransform = eigenSolver.eigenvectors() * eigenSolver.eigenvalues().cwiseSqrt().asDiagonal();
    *used_eigen = 1;
  }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) A[(size_t)i * d + j] = normTransform(i, j);
}

// a4 + a5 as the reference runs them: status 0 = Cholesky branch, 1 = eigen branch (never fails: a matrix with a
// negative eigenvalue yields NaN samples and a NaN price in the reference, and so here).
ORACLE_API double oracle_mc_basket(double S0, double E, double r, double sigma, double T,
                                   long long N, int cp, int d, double rho, const double* Z,
                                   double* sum_out, double* sumsq_out, int* status) {
  std::vector<double> A((size_t)d * d);
  oracle_mvn_transform(d, rho, A.data(), status);
  if (*status == 0) return basket_core(S0, E, r, sigma, T, N, cp, d, A.data(), Z, sum_out, sumsq_out, false);
  std::vector<double> S0v((size_t)d, S0), sg((size_t)d, sigma), w((size_t)d, 1.0 / (double)d);
  return oracle_mc_basket_general(S0v.data(), E, r, sg.data(), T, N, cp, d, A.data(), w.data(), Z, sum_out, sumsq_out);
}

// the same pricing loop (mc_eur_multi.cpp:23-34) with what the reference hard-wires made explicit:
// a FULL d x d normal transform A (mvn.h:66-76: the Cholesky factor, or eigenvectors * sqrt(eigenvalues) when the
// matrix is only positive semi-definite), per-asset spot, volatility and weight. Bt = A * Z, Z[n*d + a].
ORACLE_API double oracle_mc_basket_general(const double* S0, double E, double r, const double* sigma, double T,
                                           long long N, int cp, int d, const double* A, const double* w,
                                           const double* Z, double* sum_out, double* sumsq_out) {
  double acc = 0, acc2 = 0;
  std::vector<double> bt((size_t)d);
  for (long long n = 0; n < N; ++n) {
    const double* z = Z + n * (long long)d;
    for (int a = 0; a < d; ++a) {
      double s = 0;
      for (int k = 0; k < d; ++k) s += A[(size_t)a * d + k] * z[k];
      bt[a] = s;
    }
    double basket = 0;
    for (int a = 0; a < d; ++a)
      basket += w[a] * S0[a] * std::exp((r - sigma[a] * sigma[a] / 2) * T + sigma[a] * bt[a]);
    double v = payoff(basket, E, cp);
    acc += v;
    acc2 += v * v;
  }
  if (sum_out) *sum_out = acc;
  if (sumsq_out) *sumsq_out = acc2;
  return (std::exp(-r * T) * acc) / (double)N;
}

// Timing-only twin with the OpenMP placement of reference src/mc_eur_multi_omp.cpp:31-46: the
// sample generation stays serial (mvnorm is called outside the parallel region), the payoff loop is
// `omp for schedule(dynamic,1000) reduction(+)`. Draws its own mt19937 normals (d*N of them).
ORACLE_API double oracle_mc_basket_omp_timed(double S0, double E, double r, double sigma, double T,
                                             long long N, int cp, int d, double rho, uint64_t seed,
                                             int threads, double* seconds) {
  omp_set_num_threads(threads);
  double t0 = omp_get_wtime();
  std::vector<double> L((size_t)d * d);
  if (oracle_chol_equicorr(d, rho, L.data())) return NAN;
  std::vector<double> Z((size_t)d * (size_t)N);
  oracle_normals_mt19937(seed, 1.0, (long long)d * N, Z.data());
  double res = basket_core(S0, E, r, sigma, T, N, cp, d, L.data(), Z.data(), nullptr, nullptr, true);
  *seconds = omp_get_wtime() - t0;
  return res;
}

// ---------------------------------------------------------------------------------------------
// a9: reference src/binom_embar.cpp:5-50 and comb(), include/common.h:63-72 (O(N) per call, so the
// whole sum is O(N^2) exactly like the reference).
namespace {
double ln_comb(int N, int i) {
  if (i == 0 || i == N) return 0;
  if (i == 1 || i == (N - 1)) return std::log(N);
  double s = 0;
  for (int n = N; n > i; --n) s += std::log((double)n);
  for (int j = 2; j <= (N - i); ++j) s -= std::log((double)j);
  return s;
}
}  // namespace

// Lattice parameters exactly as the reference derives them (binom_embar.cpp:19-27): the product
// must reuse these expressions on the host because sqrt(beta^2-1) cancels catastrophically [F5].
ORACLE_API void oracle_binom_params(double r, double sigma, double T, int N, double* u, double* d,
                                    double* p, double* q) {
  double dt = (double)T / (double)N;
  double beta = 0.5 * (std::exp(-r * dt) + std::exp((r + sigma * sigma) * dt));
  *u = beta + std::sqrt(beta * beta - 1);
  *d = beta - std::sqrt(beta * beta - 1);
  double R = std::exp(r * dt);
  *p = (R - *d) / (*u - *d);
  *q = 1 - *p;
}

ORACLE_API double oracle_binom(double S0, double E, double r, double sigma, double T, int N, int cp,
                               int threads) {
  double u, d, p, q;
  oracle_binom_params(r, sigma, T, N, &u, &d, &p, &q);
  const int until = (N % 2 != 0) ? (N + 1) / 2 : N / 2;
  double result = 0;
  if (threads > 1) omp_set_num_threads(threads);
  // serial order when threads <= 1 (bit-faithful to binom_embar.cpp:34-46); the threaded variant
  // mirrors binom_embar_omp.cpp's reduction and is used only for timing.
#pragma omp parallel for schedule(dynamic, 1000) reduction(+ : result) if (threads > 1)
  for (int i = 0; i < until; ++i) {
    double c = ln_comb(N, i);
    double b1 = c + i * std::log(p) + (N - i) * std::log(q);
    double b2 = c + (N - i) * std::log(p) + i * std::log(q);
    result += std::exp(b1) * payoff(S0 * std::pow(u, i) * std::pow(d, N - i), E, cp);
    result += std::exp(b2) * payoff(S0 * std::pow(u, N - i) * std::pow(d, i), E, cp);
    if (i == 0 && N % 2 == 0) {
      double bm = ln_comb(N, N / 2) + N / 2 * std::log(p) + N / 2 * std::log(q);
      result += std::exp(bm) * payoff(S0 * std::pow(u, N / 2) * std::pow(d, N / 2), E, cp);
    }
  }
  return std::exp(-r * T) * result;
}

// Backward-induction trees: reference src/binom_vanilla_eur.cpp:15-41 and
// src/binom_vanilla_amer.cpp:15-42. Same lattice parameters as binom_embar; terminal layer
// v[i] = payoff(S0 u^i d^(N-i)) (eur: max((S-E)*cp, 0) with cp a double, :30; amer: payoff(), :29), then for
// n = N-1..0, i = 0..n:  v[i] = (p v[i+1] + q v[i]) / R   (eur :35), and for the American tree the max with the
// immediate payoff payoff(S0 pow(u,i) pow(d,n-i)) (amer :33-35). In place, ascending i, so v[i+1] is still the
// layer-(n+1) value when v[i] is overwritten.
ORACLE_API double oracle_binom_tree(double S0, double E, double r, double sigma, double T, int N, int cp,
                                    int american) {
  double u, d, p, q;
  oracle_binom_params(r, sigma, T, N, &u, &d, &p, &q);
  const double dt = (double)T / (double)N;
  const double R = std::exp(r * dt);
  std::vector<double> v(N + 1);
  for (int i = 0; i <= N; ++i) {
    const double S = S0 * std::pow(u, i) * std::pow(d, N - i);
    v[i] = american ? payoff(S, E, cp) : std::max((S - E) * (double)cp, (double)0.0);
  }
  for (int n = N - 1; n >= 0; --n)
    for (int i = 0; i <= n; ++i) {
      const double jatk = (p * v[i + 1] + q * v[i]) / R;
      if (american) {
        const double sij = payoff(S0 * std::pow(u, i) * std::pow(d, n - i), E, cp);
        v[i] = std::max(jatk, sij);
      } else {
        v[i] = jatk;
      }
    }
  return v[0];
}

// ---------------------------------------------------------------------------------------------
// CPU restatement of the PRODUCT's counter-based normal stream (include/pcf.h "normal stream v1"),
// used to check the CUDA generator: Philox4x32-10 (Salmon et al., SC'11; Random123 KAT vectors in
// tests/test_oracle.py) keyed by the 64-bit seed, counter = (index lo, index hi, draw/2, stream);
// one call yields a Box-Muller pair:
//   X1 = x1:x0, X2 = x3:x2
//   u1 = 1 - (X1 >> 12)*2^-52 in (0,1],  u2 = ((X2 >> 6) + 1/2)*2^-58 in (0,1)
//   (z_even, z_odd) = sqrt(-2 ln u1) * (cos 2 pi u2, sin 2 pi u2), evaluated here in long double
ORACLE_API void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

ORACLE_API void oracle_normal_pair(uint64_t seed, uint64_t index, uint32_t block, uint32_t stream,
                                   double* z_even, double* z_odd) {
  uint32_t ctr[4] = {(uint32_t)index, (uint32_t)(index >> 32), block, stream};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t x[4];
  oracle_philox4x32_10(ctr, key, x);
  const uint64_t X1 = ((uint64_t)x[1] << 32) | x[0], X2 = ((uint64_t)x[3] << 32) | x[2];
  const double u1 = 1.0 - (double)(X1 >> 12) * 0x1p-52;                 // exact
  const long double u2 = ((long double)(X2 >> 6) + 0.5L) * 0x1p-58L;      // exact in the x87 64-bit mantissa
  const long double R = sqrtl(-2.0L * logl((long double)u1));
  const long double two_pi = 6.283185307179586476925286766559005768L;
  *z_even = (double)(R * cosl(two_pi * u2));
  *z_odd = (double)(R * sinl(two_pi * u2));
}

// out[i*T + t] = scale * z(index0 + i, t), t < T  (the layout every replay stream above uses).
ORACLE_API void oracle_normal_stream(uint64_t seed, uint32_t stream, uint64_t index0, long long count,
                                     int T, double scale, double* out) {
  for (long long i = 0; i < count; ++i)
    for (int t = 0; t < T; t += 2) {
      double ze, zo;
      oracle_normal_pair(seed, index0 + (uint64_t)i, (uint32_t)(t >> 1), stream, &ze, &zo);
      out[i * (long long)T + t] = scale * ze;
      if (t + 1 < T) out[i * (long long)T + t + 1] = scale * zo;
    }
}

// ---------------------------------------------------------------------------------------------
// Timing-only OpenMP twins ("port" CPU baseline, used when oracle/_ref is not available). Same
// structure as reference src/mc_asia_omp.cpp:18-47 / src/mc_eur_omp.cpp:16-33: one mt19937 +
// normal_distribution per thread seeded seed*(1+thread), `omp for schedule(dynamic,1000)
// reduction(+)`. Returns the price; *seconds = wall time of the pricing function.
ORACLE_API double oracle_mc_asia_omp_timed(double S0, double E, double r, double sigma, double T,
                                           long long N, int M, int cp, uint64_t seed, int threads,
                                           double* seconds) {
  omp_set_num_threads(threads);
  double t0 = omp_get_wtime();
  double result = 0;
#pragma omp parallel
  {
    const double dt = (double)T / (double)M;
    std::mt19937 gen;
    gen.seed((std::mt19937::result_type)(seed * (1 + omp_get_thread_num())));
    std::normal_distribution<> norm{0, std::sqrt(dt)};
#pragma omp for nowait schedule(dynamic, 1000) reduction(+ : result)
    for (long long n = 0; n < N; ++n) {
      double St = S0, I = 0;
      for (int m = 0; m < M; ++m) {
        double dBi = norm(gen);
        I += St * (1 + r * dt / 2 + sigma * dBi / 2);
        St *= std::exp((r - sigma * sigma / 2) * dt + sigma * dBi);
      }
      result += payoff(I / (double)M, E, cp);
    }
  }
  *seconds = omp_get_wtime() - t0;
  return (std::exp(-r * T) * result) / (double)N;
}

ORACLE_API double oracle_mc_eur_omp_timed(double S0, double E, double r, double sigma, double T,
                                          long long N, int cp, uint64_t seed, int threads,
                                          double* seconds) {
  omp_set_num_threads(threads);
  double t0 = omp_get_wtime();
  double result = 0;
#pragma omp parallel
  {
    std::mt19937 gen;
    gen.seed((std::mt19937::result_type)(seed * (1 + omp_get_thread_num())));
    std::normal_distribution<> norm{0, std::sqrt(T)};
#pragma omp for nowait schedule(dynamic, 1000) reduction(+ : result)
    for (long long n = 0; n < N; ++n)
      result += payoff(S0 * std::exp((r - sigma * sigma / 2) * T + sigma * norm(gen)), E, cp);
  }
  *seconds = omp_get_wtime() - t0;
  return (std::exp(-r * T) * result) / (double)N;
}
