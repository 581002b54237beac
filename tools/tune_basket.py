"""Dev tool: general basket kernel, d = 16 dense covariance: launch shapes x guarded / guard-free body."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
rng = np.random.default_rng(16)
B = rng.standard_normal((16, 16))
S0, sig, w = rng.uniform(80, 120, 16), rng.uniform(.1, .4, 16), rng.dirichlet(np.ones(16))
cov = B @ B.T / 16 + .2 * np.eye(16)
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2 * 10**8
for guarded in ("1", None):
    os.environ.pop("PCF_BASKET_GUARDED", None)
    if guarded:
        os.environ["PCF_BASKET_GUARDED"] = guarded
    for shape in ("13", "12"):
        os.environ["PCF_BASKET_GEN"] = shape
        best = None
        for i in range(3):
            r = pcf.mc_basket(S0, 100., .05, sig, 1., N, "call", 16, weights=w, cov=cov, seed=1)
            best = r if best is None or r.seconds_kernel < best.seconds_kernel else best
        print(f"general basket d=16 N={N} {'guarded' if guarded else 'one block'} shape {shape}: {best.seconds_kernel*1e3:.3f} ms "
              f"{best.units/best.seconds_kernel:.4e} paths/s price {best.price!r}", flush=True)
# the reference's basket through the general kernel (PCF_BASKET_GENERAL) for comparison with the fast path
os.environ.pop("PCF_BASKET_GUARDED", None)
os.environ["PCF_BASKET_GEN"] = "13"
pcf.shutdown()
