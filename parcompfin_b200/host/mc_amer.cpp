// bin/mc_amer <call|put> S0 E r sigma T N M [gpus] -- drop-in for reference src/mc_amer.cpp:116-156
#include "frontend.h"
int main(int argc, char* argv[]) {
  frontend::Clock overall;
  frontend::need_args(argc, 9, "mc_amer <call|put> S0 E r sigma T N M [gpus]");
  std::string payoff_fun = argv[1];
  pcf_params p = frontend::base_params(payoff_fun, argv);
  p.M = frontend::getArg(argv, 8);
  if (std::getenv("PCF_AMER_LSM")) p.flags |= PCF_FLAG_AMER_LSM;  // textbook exercise rule (off by default)
  int gpus = argc > 9 ? frontend::getArg(argv, 9) : 0;
  return frontend::run("mc_amer", pcf_mc_amer, p, payoff_fun, gpus, overall, p.M, 1);
}
