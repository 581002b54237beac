"""Dev tool: one small call of every kernel family, sized for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import parcompfin_b200 as pcf
pcf.init(1)
P = (100., 100., .05, .2, 1.)
print("mc_eur", pcf.mc_eur(*P, 200_003, "call", seed=1).price)
print("mc_asia", pcf.mc_asia(*P, 20_001, 13, "call", seed=1).price)
print("basket equi", pcf.mc_eur_multi(*P, 20_001, "call", 16, .5, seed=1).price)
print("basket general", pcf.mc_basket([90., 100., 110.], 100, .05, [.1, .2, .3], 1, 20_001, "put", 3,
                                      cov=np.array([[1, .2, .1], [.2, 1, .3], [.1, .3, 1.]]), seed=1).price)
print("basket eigen", pcf.mc_basket(100, 100, .05, .2, 1, 20_001, "call", 4, cov=np.ones((4, 4)), seed=1).price)
for N, M in ((30_002, 7), (4_098, 50), (20_000, 130)):   # ragged tails, uint8 and uint16 date arrays
    print("mc_amer", N, M, pcf.mc_amer(*P, N, M, "put", seed=1).price, pcf.mc_amer(*P, N, M, "put", seed=1, lsm=True).price)
rng = np.random.default_rng(0)
print("mc_amer replay", pcf.mc_amer(*P, 10_000, 20, "call", replay=rng.standard_normal(5_000 * 20) * (1 / 20) ** .5).price)
print("mc_asia replay", pcf.mc_asia(*P, 5_000, 20, "call", replay=rng.standard_normal(5_000 * 20) * (1 / 20) ** .5).price)
for N in (1, 2, 1001, 300_000):
    print("binom", N, pcf.binom(*P, N, "call").price, pcf.binom(*P, N, "put", screen=False).price)
for N in (1, 63, 64, 65, 1000, 5000):
    print("trees", N, pcf.binom_vanilla_eur(*P, N, "put").price, pcf.binom_vanilla_amer(*P, N, "put").price)
for shape in ("1216", "1312", "1404", "14044", "1604", "48"):  # pinned CTA shapes (halo exchange), the warp kernel
    os.environ["PCF_TREE"] = shape
    print("trees", shape, pcf.binom_vanilla_eur(*P, 3001, "put").price, pcf.binom_vanilla_amer(*P, 3001, "put").price)
os.environ.pop("PCF_TREE")
print("basket general d=16 (guard-free body)", pcf.mc_basket(100, 100, .05, .2, 1, 20_001, "call", 16,
                                                           cov=.5 * np.eye(16) + .5, seed=1).price)
print("stream", pcf.normal_stream(3, 1, 0, 100, 5).sum(), pcf.philox4x32_10((0, 0, 0, 0), (0, 0)))
pcf.shutdown()
print("done")
