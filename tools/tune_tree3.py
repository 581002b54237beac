"""Dev tool: warp-trapezoid vs CTA-cooperative tree kernels (PCF_TREE shapes; auto = per-launch choice), N = 1e4 .. 4e5."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
P = (100., 100., .05, .2, 1.)
shapes = sys.argv[1:] or ["auto", "44", "1208", "1312", "1404", "1604", "1804", "1408", "1608", "1808", "1812"]
ref = {}
for N in (10_000, 100_000, 400_000, 1_000_000):
    for shape in shapes:
        os.environ.pop("PCF_TREE", None)
        if shape != "auto":
            os.environ["PCF_TREE"] = shape
        for fn, name in ((pcf.binom_vanilla_eur, "eur"), (pcf.binom_vanilla_amer, "amer")):
            best = None
            for i in range(3 if N <= 100_000 else 2):
                r = fn(*P, N, "put")
                if best is None or r.seconds_kernel < best.seconds_kernel:
                    best = r
            same = ref.setdefault((N, name), best.price) == best.price
            print(f"N={N} shape {shape} {name}: kernel {best.seconds_kernel*1e3:.3f} ms  call {best.seconds_total*1e3:.3f} ms  "
                  f"{best.units/best.seconds_kernel:.3e} nodes/s launches {best.launches} price {best.price!r} "
                  f"{'same' if same else 'DIFFERENT'}", flush=True)
pcf.shutdown()
