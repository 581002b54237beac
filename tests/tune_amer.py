"""Dev tool: sweep PCF_AMER_GEN (pairs per thread x CTAs per SM) of the American path kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
a = (100., 100., .05, .2, 1.)
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**8
for v in "14,22,23,32,41,42,61".split(","):
    os.environ["PCF_AMER_GEN"] = v
    best = 1e9
    for i in range(3):
        r = pcf.mc_amer(*a, N, 50, "put", seed=1)
        best = min(best, r.seconds_kernel)
    print(f"variant {v}: {best*1e3:.3f} ms  price {r.price!r}", flush=True)
pcf.shutdown()
