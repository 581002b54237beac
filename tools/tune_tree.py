"""Dev tool: launch shapes of the tree kernel (PCF_TREE) at N = 1e5 and beyond."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
P = (100., 100., .05, .2, 1.)
for N in (100_000, 400_000):
    for shape in ("22", "44", "48", "84", "88"):
        os.environ["PCF_TREE"] = shape
        for fn, name in ((pcf.binom_vanilla_eur, "eur"), (pcf.binom_vanilla_amer, "amer")):
            best = None
            for i in range(2):
                r = fn(*P, N, "put")
                if best is None or r.seconds_kernel < best.seconds_kernel:
                    best = r
            print(f"N={N} shape {shape} {name}: kernel {best.seconds_kernel*1e3:.3f} ms  call {best.seconds_total*1e3:.3f} ms  "
                  f"{best.units/best.seconds_kernel:.3e} nodes/s launches {best.launches} price {best.price!r}", flush=True)
pcf.shutdown()
