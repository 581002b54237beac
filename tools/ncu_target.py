"""Profiling target (dev tool): `python tools/ncu_target.py <workload> <N> [reps]` runs one workload through
the C ABI so that ncu can capture its kernels. Not a test."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
name, N = sys.argv[1], int(float(sys.argv[2]))
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pcf.init(1)
a = (100., 100., .05, .2, 1.)
for i in range(reps):
    if name == "mc_asia": r = pcf.mc_asia(*a, N, 252, "call", seed=i)
    elif name == "mc_eur": r = pcf.mc_eur(*a, N, "call", seed=i)
    elif name == "mc_eur_multi": r = pcf.mc_eur_multi(*a, N, "call", 16, .5, seed=i)
    elif name == "mc_basket_general":
        import numpy as np
        rng = np.random.default_rng(16)
        B = rng.standard_normal((16, 16))
        r = pcf.mc_basket(rng.uniform(80, 120, 16), 100., .05, rng.uniform(.1, .4, 16), 1., N, "call", 16,
                          weights=rng.dirichlet(np.ones(16)), cov=B @ B.T / 16 + .2 * np.eye(16), seed=i)
    elif name == "mc_amer": r = pcf.mc_amer(*a, N, 50, "put", seed=i)
    elif name == "binom_embar": r = pcf.binom(*a, N, "call")
    elif name == "binom_embar_noscreen": r = pcf.binom(*a, N, "call", screen=False)
    elif name == "tree_eur": r = pcf.binom_vanilla_eur(*a, N, "put")
    elif name == "tree_amer": r = pcf.binom_vanilla_amer(*a, N, "put")
    print(name, N, r.price, r.seconds_kernel, r.units / r.seconds_kernel)
pcf.shutdown()
