"""Dev tool: timing experiments on the American path kernel (PCF_AMER_DBG=16 stores only the last row)."""
import os, sys
sys.path.insert(0, os.getcwd())
import parcompfin_b200 as pcf
pcf.init(1)
a = (100., 100., .05, .2, 1.)
N = 10**8
for rep in range(2):
    for gen in ("22", "23", "42", "41", "14"):
        os.environ["PCF_AMER_GEN"] = gen
        for dbg in (0, 16):
            os.environ["PCF_AMER_DBG"] = str(dbg)
            r = pcf.mc_amer(*a, N, 50, "put", seed=1)
            print(f"gen {gen} dbg {dbg}: {r.seconds_kernel*1e3:.3f} ms price {r.price}", flush=True)
pcf.shutdown()
