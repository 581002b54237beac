import os, sys
sys.path.insert(0, os.getcwd())
import parcompfin_b200 as pcf
pcf.init(1)
a = (100., 100., .05, .2, 1.)
N = 10**8
for shape in ("1,256,2", "2,128,3"):
    os.environ["PCF_AMER_SWEEP"] = shape
    for dbg in (0, 1, 2, 3):
        os.environ["PCF_AMER_DBG"] = str(dbg)
        best = 1e9
        for i in range(2):
            r = pcf.mc_amer(*a, N, 50, "put", seed=1)
            best = min(best, r.seconds_kernel)
        print(f"shape {shape} dbg {dbg}: {best*1e3:.3f} ms price {r.price}", flush=True)
pcf.shutdown()
