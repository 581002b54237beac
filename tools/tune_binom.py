"""Dev tool: binomial term kernel with and without the screening pass; tree launch chain with and without PDL."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
P = (100., 100., .05, .2, 1.)
for N in (10**5, 10**7, 10**8, 2**31 - 1):
    for screen in (True, False):
        best = None
        for i in range(4):
            r = pcf.binom(*P, N, "call", screen=screen)
            if best is None or r.seconds_kernel < best.seconds_kernel:
                best = r
        print(f"binom N={N} screen={screen}: kernel {best.seconds_kernel*1e3:.4f} ms call {best.seconds_total*1e3:.4f} ms "
              f"{best.units/best.seconds_kernel:.3e} terms/s price {best.price!r}", flush=True)
for N in (10_000, 100_000, 400_000):
    for pdl in ("1", "0"):
        os.environ["PCF_TREE_PDL"] = pdl
        for shape in ("44", "48", "68", "88"):
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
