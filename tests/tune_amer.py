"""Dev tool: sweep the launch shapes of the American kernels (PCF_AMER_GEN, PCF_AMER_SWEEP)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
a = (100., 100., .05, .2, 1.)
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**8
what = sys.argv[2] if len(sys.argv) > 2 else "both"


def run(tag):
    best = 1e9
    for i in range(3):
        r = pcf.mc_amer(*a, N, 50, "put", seed=1)
        best = min(best, r.seconds_kernel)
    print(f"{tag}: {best*1e3:.3f} ms  price {r.price!r}", flush=True)


if what in ("gen", "both"):
    for v in "14,22,23,32,41,42,61".split(","):
        os.environ["PCF_AMER_GEN"] = v
        run(f"gen {v}")
    os.environ.pop("PCF_AMER_GEN")
if what in ("sweep", "both"):
    for u in (1, 2, 4):
        for blk, per in ((128, 2), (128, 3), (128, 4), (128, 6), (128, 8), (256, 1), (256, 2), (256, 3), (256, 4)):
            os.environ["PCF_AMER_SWEEP"] = f"{u},{blk},{per}"
            run(f"sweep {u},{blk},{per}")
pcf.shutdown()
