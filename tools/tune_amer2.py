"""Dev tool: American sweep A/B -- alternating tile direction + L2 keep policy (PCF_AMER_DBG=32 turns it off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
a = (100., 100., .05, .2, 1.)
N = 10**8
best = {}
for rep in range(4):
    for dbg in ("0", "32"):
        os.environ["PCF_AMER_DBG"] = dbg
        r = pcf.mc_amer(*a, N, 50, "put", seed=1)
        best[dbg] = min(best.get(dbg, 1e9), r.seconds_kernel)
        print(f"rep {rep} dbg={dbg}: {r.seconds_kernel*1e3:.3f} ms price {r.price!r}", flush=True)
print({k: round(v * 1e3, 3) for k, v in best.items()})
pcf.shutdown()
