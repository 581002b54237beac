// bin/mc_asia <call|put> S0 E r sigma T N M [gpus] -- drop-in for reference src/mc_asia.cpp:42-81
#include "frontend.h"
int main(int argc, char* argv[]) {
  frontend::Clock overall;
  frontend::need_args(argc, 9, "mc_asia <call|put> S0 E r sigma T N M [gpus]");
  std::string payoff_fun = argv[1];
  pcf_params p = frontend::base_params(payoff_fun, argv);
  p.M = frontend::getArg(argv, 8);
  int gpus = argc > 9 ? frontend::getArg(argv, 9) : 0;
  return frontend::run("mc_asia", pcf_mc_asia, p, payoff_fun, gpus, overall, p.M, 1);
}
