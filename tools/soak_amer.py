"""Soak test of the American sweep (dev tool): the same call repeated many times must return the same bits -- a race in
the TMA ring, the deferred gathers or the date chain would show as run-to-run noise (it did once, profiles/r1_notes.md).
Also interleaves different shapes between repetitions so that stale shared/global state would be noticed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
P = (100., 100., .05, .2, 1.)
cases = [("put 1e8 x 50", lambda: pcf.mc_amer(*P, 100_000_000, 50, "put", seed=7), 12),
         ("put lsm 2e7 x 50", lambda: pcf.mc_amer(*P, 20_000_000, 50, "put", seed=7, lsm=True), 12),
         ("call 1e7 x 130 (uint16 dates)", lambda: pcf.mc_amer(100, 110, .02, .75, 1, 10_000_000, 130, "call", seed=7), 12),
         ("call lsm 1e7 x 20", lambda: pcf.mc_amer(100, 110, .02, .75, 1, 10_000_000, 20, "call", seed=7, lsm=True), 12),
         ("put 4098 x 7", lambda: pcf.mc_amer(*P, 4098, 7, "put", seed=7), 50)]
ref = {}
bad = 0
for rep in range(max(c[2] for c in cases)):
    for name, fn, n in cases:
        if rep >= n:
            continue
        r = fn()
        key = (r.price, r.sum, r.sumsq)
        if name not in ref:
            ref[name] = key
            print(f"{name}: price {r.price!r} sum {r.sum!r} sumsq {r.sumsq!r} {r.seconds_kernel*1e3:.3f} ms", flush=True)
        elif key != ref[name]:
            bad += 1
            print(f"MISMATCH {name} rep {rep}: {key} vs {ref[name]}", flush=True)
print("soak:", "FAILED" if bad else "ok", f"({sum(c[2] for c in cases)} calls, {bad} mismatches)")
pcf.shutdown()
sys.exit(1 if bad else 0)
