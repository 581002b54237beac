"""Dev tool: sweep the launch shapes of the American path kernel (PCF_AMER_GEN) and the ring depth of the sweep
(PCF_AMER_SWEEP=<stages>); the sweep shapes themselves are in tools/tune_amer_chain.py."""
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
    shapes = ((2, 3, 0), (2, 3, 1), (3, 3, 0), (2, 2, 0), (2, 2, 1), (3, 2, 1), (4, 2, 0))
    best = {k: 1e9 for k in shapes}
    for rep in range(4):  # interleaved: the GPU's clocks drift as it warms up, so no shape is always first
        for stages, per, early in shapes:
            os.environ["PCF_AMER_SWEEP"] = f"{stages},{per},{early}"
            r = pcf.mc_amer(*a, N, 50, "put", seed=1)
            best[(stages, per, early)] = min(best[(stages, per, early)], r.seconds_kernel)
            print(f"rep {rep} sweep stages={stages} ctas/sm={per} early={early}: {r.seconds_kernel*1e3:.3f} ms  price {r.price!r}", flush=True)
    for k, v in best.items():
        print(f"best stages={k[0]} ctas/sm={k[1]} early={k[2]}: {v*1e3:.3f} ms")
pcf.shutdown()
