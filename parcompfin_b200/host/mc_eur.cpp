// bin/mc_eur <call|put> S0 E r sigma T N [gpus] -- drop-in for reference src/mc_eur.cpp:29-64
#include "frontend.h"
int main(int argc, char* argv[]) {
  frontend::Clock overall;
  frontend::need_args(argc, 8, "mc_eur <call|put> S0 E r sigma T N [gpus]");
  std::string payoff_fun = argv[1];
  pcf_params p = frontend::base_params(payoff_fun, argv);
  int gpus = argc > 8 ? frontend::getArg(argv, 8) : 0;
  return frontend::run("mc_eur", pcf_mc_eur, p, payoff_fun, gpus, overall, 0, 1);
}
