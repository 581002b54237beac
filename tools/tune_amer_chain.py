"""A/B of the per-date American sweep chain (needs `make lib TUNING=1`; run with PCF_LIB=parcompfin_b200/libpcf_tuning.so):
fixed tile assignment vs tiles handed out on demand, with and without programmatic dependent launch, against the
persistent all-dates kernel; then the per-CTA tile-loop times of one date (date M/2) grouped by SM."""
import ctypes, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
P1 = (100, 100, 0.05, 0.2, 1)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000

pcf.init(1)
lib = pcf.load_library()


def per_cta(tag):
    buf = (ctypes.c_ulonglong * (3 * 1024))()
    if lib.pcf_debug_sweep(buf, 3 * 1024) != 0:
        return
    rows = [(buf[3 * i], buf[3 * i + 1] / 1965.0, buf[3 * i + 2]) for i in range(444) if buf[3 * i + 2]]
    if not rows:
        return
    us = sorted(r[1] for r in rows)
    tl = sorted(r[2] for r in rows)
    print(f"   [{tag}] date M/2, {len(rows)} CTAs: consumer loop us min {us[0]:.0f} / median {us[len(us)//2]:.0f} / max {us[-1]:.0f}; "
          f"tiles per CTA min {tl[0]} / median {tl[len(tl)//2]} / max {tl[-1]}")
    by_sm = collections.defaultdict(list)
    for sm, t, n in rows:
        by_sm[sm].append(t / max(n, 1))
    per = sorted((sum(v) / len(v), sm) for sm, v in by_sm.items())
    print("   us per tile by SM: fastest", " ".join(f"{sm}:{t:.2f}" for t, sm in per[:4]), "| slowest",
          " ".join(f"{sm}:{t:.2f}" for t, sm in per[-4:]))


VARIANTS = [("8x3, no PDL (round-1 shape)", {"PCF_AMER_SHAPE": "83", "PCF_AMER_NOPDL": "1"}),
            ("8x3", {"PCF_AMER_SHAPE": "83"}),
            ("8x3 on demand", {"PCF_AMER_SHAPE": "83", "PCF_AMER_ONDEMAND": "1"}),
            ("7x3", {"PCF_AMER_SHAPE": "73"}),
            ("11x2", {"PCF_AMER_SHAPE": "112"}),
            ("11x2, 3 stages", {"PCF_AMER_SHAPE": "112", "PCF_AMER_SWEEP": "3"}),
            ("20x1, 3 stages", {"PCF_AMER_SHAPE": "201", "PCF_AMER_SWEEP": "3"}),
            ("23x1, 3 stages (default)", {}),
            ("23x1, 3 stages, no PDL", {"PCF_AMER_NOPDL": "1"}),
            ("23x1, 2 stages", {"PCF_AMER_SWEEP": "2"}),
            ("23x1, 4 stages", {"PCF_AMER_SWEEP": "4"}),
            ("23x1 on demand, 3 stages", {"PCF_AMER_SWEEP": "3", "PCF_AMER_ONDEMAND": "1"}),
            ("24x1, 3 stages", {"PCF_AMER_SHAPE": "241", "PCF_AMER_SWEEP": "3"}),
            ("ENVELOPE (wrong results) 23x1, no gathers", {"PCF_AMER_ENVELOPE": "1"}),
            ("ENVELOPE 23x1 3 stages, no gathers", {"PCF_AMER_ENVELOPE": "1", "PCF_AMER_SWEEP": "3"}),
            ("ENVELOPE 23x1 3 stages, gathers issued but not waited for", {"PCF_AMER_ENVELOPE": "8", "PCF_AMER_SWEEP": "3"}),
            ("ENVELOPE 23x1 3 stages, gathers redirected to row m", {"PCF_AMER_ENVELOPE": "16", "PCF_AMER_SWEEP": "3"}),
            ("ENVELOPE 23x1, no date stores", {"PCF_AMER_ENVELOPE": "4"}),
            ("ENVELOPE 23x1, ring only", {"PCF_AMER_ENVELOPE": "2"})]
KNOBS = ("PCF_AMER_SHAPE", "PCF_AMER_ONDEMAND", "PCF_AMER_NOPDL", "PCF_AMER_SWEEP", "PCF_AMER_ENVELOPE")
for rep in range(3):
    for name, env in VARIANTS:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        r = pcf.mc_amer(*P1, N, 50, "put", seed=20240229)
        print(f"rep {rep} {name}: {r.seconds_kernel*1e3:.3f} ms  launches {r.launches}  price {r.price!r}", flush=True)
        if rep == 2:
            per_cta(name)
pcf.shutdown()
