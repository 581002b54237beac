"""A/B of the binomial term kernel's launch shape (needs `make lib TUNING=1`; run with
PCF_LIB=parcompfin_b200/libpcf_tuning.so): PCF_BINOM_VARIANT = <pairs screened side by side><CTAs per SM>."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
P = (100., 100., .05, .2, 1.)
for N in (10**7, 10**8, 2**31 - 1):
    for screen in (True, False):
        for var in ("14", "13", "24", "23", "44", "43", "42", "82"):
            os.environ["PCF_BINOM_VARIANT"] = var
            best = None
            for i in range(5):
                r = pcf.binom(*P, N, "call", screen=screen)
                if best is None or r.seconds_kernel < best.seconds_kernel:
                    best = r
            print(f"binom N={N} screen={screen} variant {var}: kernel {best.seconds_kernel*1e3:.4f} ms "
                  f"{best.units/best.seconds_kernel:.3e} terms/s price {best.price!r}", flush=True)
pcf.shutdown()
