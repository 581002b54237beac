// bin/binom_vanilla_eur <call|put> S0 E r sigma T N -- drop-in for reference src/binom_vanilla_eur.cpp:44-80
// (the tree the run-scripts use to bake `comparison`, runscript_mc_eur.sh:23). Method column: "CUDA_vanilla"
// where the reference prints "Serial_vanilla".
#include "frontend.h"
int main(int argc, char* argv[]) {
  frontend::Clock overall;
  frontend::need_args(argc, 8, "binom_vanilla_eur <call|put> S0 E r sigma T N");
  std::string payoff_fun = argv[1];
  pcf_params p = frontend::base_params(payoff_fun, argv);
  return frontend::run("binom_vanilla_eur", pcf_binom_vanilla_eur, p, payoff_fun, 1, overall, 0, 1, "CUDA_vanilla");
}
