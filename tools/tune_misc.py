"""Dev tool: sweep launch shapes of the European and basket kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
a = (100., 100., .05, .2, 1.)
for v in "14,22,23,32,33,41,42,51,61,81".split(","):
    os.environ["PCF_EUR_VARIANT"] = v
    best = 0
    for i in range(3):
        r = pcf.mc_eur(*a, 10**9, "call", seed=1); best = max(best, r.units / r.seconds_kernel)
    print(f"eur variant {v}: {best:.4e} paths/s price {r.price!r}", flush=True)
for v in "14,22,23,32,41,42,61,81".split(","):
    os.environ["PCF_BASKET_VARIANT"] = v
    best = 0
    for i in range(3):
        r = pcf.mc_eur_multi(*a, 10**8, "call", 16, .5, seed=1); best = max(best, r.units / r.seconds_kernel)
    print(f"basket variant {v}: {best:.4e} paths/s price {r.price!r}", flush=True)
pcf.shutdown()
