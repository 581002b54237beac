"""A/B of the American sweep drivers (needs `make lib TUNING=1`; run with PCF_LIB=parcompfin_b200/libpcf_tuning.so):
per-date chain (round 1) against the persistent all-dates kernel with a 2- and a 3-deep TMA ring."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
P1 = (100, 100, 0.05, 0.2, 1)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
pcf.init(1)
for rep in range(3):
    for name, env in (("chain", {"PCF_AMER_CHAIN": "1"}),
                      ("persistent 2st", {"PCF_AMER_SWEEP": "2", "PCF_AMER_KNOBS": "1"}),
                      ("persistent 3st", {"PCF_AMER_SWEEP": "3", "PCF_AMER_KNOBS": "1"})):
        for k in ("PCF_AMER_CHAIN", "PCF_AMER_SWEEP", "PCF_AMER_L2", "PCF_AMER_NOCOOP", "PCF_AMER_KNOBS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        r = pcf.mc_amer(*P1, N, 50, "put", seed=20240229)
        print(f"rep {rep} {name}: {r.seconds_kernel*1e3:.3f} ms  launches {r.launches}  price {r.price!r}", flush=True)
        if "persistent" in name and rep == 2:
            import ctypes
            buf = (ctypes.c_ulonglong * 8)()
            pcf.load_library().pcf_debug_counters(buf)
            n = max(buf[7], 1)
            us = lambda cyc: cyc / 1965.0  # cycles -> us at 1965 MHz
            print(f"   per CTA (mean over {n} CTAs, us): date-barrier wait {us(buf[0]/n):.0f}, tile loops {us(buf[1]/n):.0f}, "
                  f"arrival {us(buf[2]/n):.0f}, total {us(buf[3]/n):.0f}; max wait {us(buf[4]):.0f}, max tiles {us(buf[5]):.0f}, "
                  f"max total {us(buf[6]):.0f}", flush=True)
pcf.shutdown()
