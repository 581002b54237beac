"""Dev tool: sweep PCF_ASIA_VARIANT (paths per thread x min blocks per SM) on the Asian kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
a = (100., 100., .05, .2, 1.)
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**8
for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else "13,14,15,16,22,23,24,32,41,42".split(",")):
    os.environ["PCF_ASIA_VARIANT"] = v
    best = 0
    for i in range(3):
        r = pcf.mc_asia(*a, N, 252, "call", seed=1)
        best = max(best, r.units / r.seconds_kernel)
    print(f"variant {v}: {best:.4e} path-steps/s price {r.price!r}", flush=True)
pcf.shutdown()
