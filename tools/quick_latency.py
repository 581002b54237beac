import sys, os, time
sys.path.insert(0, os.getcwd())
import parcompfin_b200 as pcf
pcf.init(1)
P1 = (100, 100, 0.05, 0.2, 1)
for i in range(6):
    r = pcf.mc_amer(*P1, 100_000_000, 50, "put", seed=20240229 + i)
    print("mc_amer 1e8x50: %.3f ms kernel, %.3f ms total, price %.12f launches %d" % (r.seconds_kernel*1e3, r.seconds_total*1e3, r.price, r.launches), flush=True)
for N in (10_000_000,):
    for i in range(4):
        r = pcf.mc_eur(*P1, N, "call", seed=i)
        print("mc_eur 1e7: %.1f us kernel, %.1f us total" % (r.seconds_kernel*1e6, r.seconds_total*1e6))
for N in (100_000, 1_000_000, 10_000_000, 100_000_000):
    for i in range(3):
        r = pcf.binom(*P1, N, "call")
    print("binom N=%d: %.1f us kernel, %.1f us total" % (N, r.seconds_kernel*1e6, r.seconds_total*1e6))
pcf.shutdown()
