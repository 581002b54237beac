// bin/mc_eur_multi <call|put> S0 E r sigma T N assets rho [gpus]
// drop-in for reference src/mc_eur_multi.cpp:37-79 (row: M = 0, Nr_of_assets = assets, :73-76)
#include "frontend.h"
int main(int argc, char* argv[]) {
  frontend::Clock overall;
  frontend::need_args(argc, 10, "mc_eur_multi <call|put> S0 E r sigma T N assets rho [gpus]");
  std::string payoff_fun = argv[1];
  pcf_params p = frontend::base_params(payoff_fun, argv);
  p.assets = frontend::getArg(argv, 8);
  p.rho = frontend::getArgD(argv, 9);
  int gpus = argc > 10 ? frontend::getArg(argv, 10) : 0;
  return frontend::run("mc_eur_multi", pcf_mc_eur_multi, p, payoff_fun, gpus, overall, 0, p.assets);
}
