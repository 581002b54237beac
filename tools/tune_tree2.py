"""Dev tool: tree launch chain -- programmatic dependent launch modes (0 off, 1 release at start, 2 release after the layers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
P = (100., 100., .05, .2, 1.)
for N in (10_000, 30_000, 100_000, 400_000):
    for pdl in ("0", "2", "1"):
        os.environ["PCF_TREE_PDL"] = pdl
        for shape in ("44", "48"):
            os.environ["PCF_TREE"] = shape
            for fn, name in ((pcf.binom_vanilla_eur, "eur"), (pcf.binom_vanilla_amer, "amer")):
                best = None
                for i in range(2):
                    r = fn(*P, N, "put")
                    if best is None or r.seconds_kernel < best.seconds_kernel:
                        best = r
                print(f"tree N={N} pdl={pdl} shape {shape} {name}: kernel {best.seconds_kernel*1e3:.3f} ms call "
                      f"{best.seconds_total*1e3:.3f} ms {best.units/best.seconds_kernel:.3e} nodes/s price {best.price!r}", flush=True)
pcf.shutdown()
