"""Profiling target (dev tool): `python tools/tree_target.py <eur|amer> <N> [reps]` runs one tree through the C ABI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
kind, N = sys.argv[1], int(float(sys.argv[2]))
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pcf.init(1)
fn = pcf.binom_vanilla_amer if kind == "amer" else pcf.binom_vanilla_eur
for i in range(reps):
    r = fn(100., 100., .05, .2, 1., N, "put")
    print(kind, N, os.environ.get("PCF_TREE"), r.price, r.seconds_kernel, r.launches)
pcf.shutdown()
