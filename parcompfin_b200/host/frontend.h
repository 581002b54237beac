// frontend.h -- what the five drop-in executables share: the reference's positional argv parsing
// (include/common.h:25-38), its "call"/"put" check (src/mc_eur.cpp:39-42), its two wall clocks
// (src/mc_eur.cpp:30,44-49) and its one-line CSV row (include/common.h:212-249), re-emitted
// byte-compatibly with Method = "CUDA" and Parallel = number of GPUs (the reference writes
// "OMP"/threads and "MPI"/ranks there).
//
// Beyond the reference's argv the executables accept ONE optional trailing integer [gpus]
// (where the _omp programs take [threads], src/mc_eur_omp.cpp:41) and read:
//   PCF_SEED        Philox key (default: time(), as the reference seeds from time())
//   PCF_COMPARISON  the value the reference bakes into include/comparison.h (default 0)
//   PCF_REPLAY      path of a raw little-endian float64 file of normals in reference draw order
//   PCF_VERBOSE     when set, a diagnostic line (std error, kernel seconds, units/s) on stderr
#pragma once
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>
#include "pcf.h"

namespace frontend {

// include/common.h:25-30
inline int getArg(char* argv[], int idx) {
  std::size_t pos;
  std::string arg = argv[idx];
  return std::stoi(arg, &pos);
}
// include/common.h:33-38
inline double getArgD(char* argv[], int idx) {
  std::string arg = argv[idx];
  return std::stod(arg);
}

inline int payoff_sign(const std::string& payoff_fun) {
  if (payoff_fun == "call") return 1;
  if (payoff_fun == "put") return -1;
  throw std::invalid_argument("Unknown payoff function");  // uncaught, like src/mc_eur.cpp:42
}

inline void need_args(int argc, int n, const char* usage) {
  if (argc < n) {  // the reference reads argv out of bounds here; fail cleanly instead
    std::fprintf(stderr, "usage: %s\n", usage);
    std::exit(2);
  }
}

struct Clock {
  std::chrono::time_point<std::chrono::system_clock> t0 = std::chrono::system_clock::now();
  double seconds() const {
    return std::chrono::duration<double>(std::chrono::system_clock::now() - t0).count();
  }
};

inline double env_double(const char* name, double dflt) {
  const char* s = std::getenv(name);
  return s ? std::atof(s) : dflt;
}

inline unsigned long long env_seed() {
  const char* s = std::getenv("PCF_SEED");
  return s ? std::strtoull(s, nullptr, 10) : (unsigned long long)std::time(nullptr);
}

inline std::vector<double> env_replay() {
  std::vector<double> v;
  const char* path = std::getenv("PCF_REPLAY");
  if (!path) return v;
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error(std::string("cannot open PCF_REPLAY file ") + path);
  std::streamsize bytes = f.tellg();
  f.seekg(0);
  v.resize((size_t)bytes / sizeof(double));
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(double)));
  return v;
}

// Maps a library status onto the reference's failure behaviour: the three conditions the reference
// throws std::invalid_argument for are rethrown (uncaught => abort, nothing on stdout); anything
// else is an environment failure reported on stderr.
inline void check(int status) {
  if (status == PCF_OK) return;
  if (status == PCF_EINVAL_PAYOFF || status == PCF_EODD_N || status == PCF_ESINGULAR)
    throw std::invalid_argument(pcf_strerror(status));
  std::fprintf(stderr, "libpcf: %s (%s)\n", pcf_strerror(status), pcf_last_error());
  std::exit(1);
}

inline pcf_params base_params(const std::string& payoff_fun, char* argv[]) {
  pcf_params p{};
  p.S0 = getArgD(argv, 2);
  p.E = getArgD(argv, 3);
  p.r = getArgD(argv, 4);
  p.sigma = getArgD(argv, 5);
  p.T = getArgD(argv, 6);
  p.N = getArg(argv, 7);
  p.cp = payoff_sign(payoff_fun);
  p.assets = 1;
  p.seed = env_seed();
  return p;
}

// include/common.h:212-249
inline void reporting(const std::string& method, const std::string& payoff_fun, double S0, double E, double r,
                      double sigma, double T, double time_overall, double time, double result,
                      double comparison, long long N, int parallel = 0, int M = 0, int assets = 1) {
  std::cout << std::setprecision(10) << method << "," << payoff_fun << "," << S0 << "," << E << "," << r
            << "," << sigma << "," << T << "," << N << "," << M << "," << parallel << "," << assets << ","
            << time_overall << "," << time << "," << result << "," << std::abs(result - comparison) << ","
            << result - comparison << std::endl;
}

inline void verbose(const char* what, const pcf_result& res) {
  if (!std::getenv("PCF_VERBOSE")) return;
  std::fprintf(stderr, "[%s] gpus=%d price=%.17g std_error=%.3g kernel_s=%.6f call_s=%.6f units=%lld units/s=%.4g launches=%d\n",
               what, res.gpus, res.price, res.std_error, res.seconds_kernel, res.seconds_total, res.units,
               res.seconds_kernel > 0 ? (double)res.units / res.seconds_kernel : 0.0, res.launches);
}

typedef int (*pcf_method)(const pcf_params*, pcf_result*);

// Runs one method the way every reference main does: parse -> start clock -> price -> report.
inline int run(const char* what, pcf_method fn, pcf_params p, const std::string& payoff_fun, int gpus,
               const Clock& overall, int M_field, int assets_field, const char* method = "CUDA") {
  std::vector<double> replay = env_replay();
  if (!replay.empty()) {
    p.replay = replay.data();
    p.replay_len = (long long)replay.size();
  }
  check(pcf_init(gpus));  // CUDA context + NCCL: part of T_overall, not of T_calculation
  Clock calc;
  pcf_result res{};
  check(fn(&p, &res));
  double t_calc = calc.seconds();
  double t_all = overall.seconds();
  reporting(method, payoff_fun, p.S0, p.E, p.r, p.sigma, p.T, t_all, t_calc, res.price,
            env_double("PCF_COMPARISON", 0.0), p.N, res.gpus, M_field, assets_field);
  verbose(what, res);
  pcf_shutdown();
  return EXIT_SUCCESS;
}

}  // namespace frontend
