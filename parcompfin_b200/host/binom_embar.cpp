// bin/binom_embar <call|put> S0 E r sigma T N [gpus] -- drop-in for reference src/binom_embar.cpp:53-91
// PCF_BINOM_WINDOW=1 enables the support-window shortcut (PCF_FLAG_BINOM_WINDOW).
#include "frontend.h"
int main(int argc, char* argv[]) {
  frontend::Clock overall;
  frontend::need_args(argc, 8, "binom_embar <call|put> S0 E r sigma T N [gpus]");
  std::string payoff_fun = argv[1];
  pcf_params p = frontend::base_params(payoff_fun, argv);
  if (std::getenv("PCF_BINOM_WINDOW")) p.flags |= PCF_FLAG_BINOM_WINDOW;
  int gpus = argc > 8 ? frontend::getArg(argv, 8) : 0;
  return frontend::run("binom_embar", pcf_binom_embar, p, payoff_fun, gpus, overall, 0, 1);
}
