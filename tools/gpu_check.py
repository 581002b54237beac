"""First-light GPU check: every method of libpcf.so against the CPU oracle, plus peaks. Dev tool."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import parcompfin_b200 as pcf
import oracle

def rel(a, b): return abs(a - b) / max(abs(b), 1e-300)

pcf.init(1)
print("device", pcf.device_info())
kat = [((0,0,0,0),(0,0),(0x6627e8d5,0xe169c58d,0xbc57ac4c,0x9b00dbd8)),
       ((0xffffffff,)*4,(0xffffffff,)*2,(0x408f276d,0x41c83b0e,0xa20bc7c6,0x6d5451fd)),
       ((0x243f6a88,0x85a308d3,0x13198a2e,0x03707344),(0xa4093822,0x299f31d0),(0xd16cfe09,0x94fdcceb,0x5001e420,0x24126ea1))]
for c,k,e in kat:
    g = pcf.philox4x32_10(c,k); o = oracle.philox4x32_10(c,k)
    print("philox KAT", g == e, o == e)
zs = pcf.normal_stream(7, 1, 1000, 4096, 5, 1.0); zo = oracle.normal_stream(7, 1, 1000, 4096, 5, 1.0)
print("normal stream max abs diff vs oracle", np.abs(zs - zo).max(), "mean", zs.mean(), "std", zs.std())

P = dict(S0=100., E=100., r=0.05, sigma=0.2, T=1.0)
# replay parity (reference stream)
N = 200000
w = oracle.normals_mt19937(42, 1.0, N)
g = pcf.mc_eur(100,100,.05,.2,1,N,"call",replay=w); o = oracle.mc_eur(100,100,.05,.2,1,N,"call",w)
print("mc_eur replay", g.price, o, rel(g.price,o))
N, M = 20000, 252
w = oracle.normals_mt19937(42, (1/M)**.5, N*M)
g = pcf.mc_asia(100,100,.05,.2,1,N,M,"call",replay=w); o = oracle.mc_asia(100,100,.05,.2,1,N,M,"call",w)
print("mc_asia replay", g.price, o, rel(g.price,o))
N, d = 50000, 16
Z = oracle.normals_mt19937(42, 1.0, N*d)
g = pcf.mc_eur_multi(100,100,.05,.2,1,N,"call",d,0.5,replay=Z); o = oracle.mc_basket(100,100,.05,.2,1,N,"call",d,0.5,Z)
print("basket replay", g.price, o, rel(g.price,o))
for cpn, args in (("put",(100,100,.05,.2,1)), ("call",(100,110,.02,.75,1))):
    N, M = 100000, 50
    w = oracle.normals_mt19937(42, (1/M)**.5, N//2*M)
    g = pcf.mc_amer(*args,N,M,cpn,replay=w); o = oracle.mc_amer(*args,N,M,cpn,w)
    print("mc_amer replay", cpn, g.price, o, rel(g.price,o), "launches", g.launches)
# native vs oracle fed the GPU's own normals
N = 100001
z = pcf.normal_stream(5, pcf.STREAM_EUR, 0, (N+1)//2, 2, 1.0).reshape(-1)[:N]
g = pcf.mc_eur(100,100,.05,.2,1,N,"put",seed=5); o = oracle.mc_eur(100,100,.05,.2,1,N,"put",z)
print("mc_eur native-vs-dump", g.price, o, rel(g.price,o))
N, M = 5000, 253
z = pcf.normal_stream(5, pcf.STREAM_ASIA, 0, N, M, (1/M)**.5)
g = pcf.mc_asia(100,100,.05,.2,1,N,M,"call",seed=5); o = oracle.mc_asia(100,100,.05,.2,1,N,M,"call",z)
print("mc_asia native-vs-dump", g.price, o, rel(g.price,o))
N, d = 20000, 16
z = pcf.normal_stream(5, pcf.STREAM_BASKET, 0, N, d, 1.0)
g = pcf.mc_eur_multi(100,100,.05,.2,1,N,"call",d,0.5,seed=5); o = oracle.mc_basket(100,100,.05,.2,1,N,"call",d,0.5,z)
print("basket native-vs-dump", g.price, o, rel(g.price,o))
N, M = 20000, 50
z = pcf.normal_stream(5, pcf.STREAM_AMER, 0, N//2, M, (1/M)**.5)
g = pcf.mc_amer(100,100,.05,.2,1,N,M,"put",seed=5); o = oracle.mc_amer(100,100,.05,.2,1,N,M,"put",z)
print("mc_amer native-vs-dump", g.price, o, rel(g.price,o))
# binomial
for N in (100, 1000, 1001, 10000):
    g = pcf.binom(100,110,.02,.75,1,N,"call"); o = oracle.binom(100,110,.02,.75,1,N,"call")
    print("binom", N, g.price, o, rel(g.price,o))
for N in (100000, 10**6, 10**7, 10**8):
    g = pcf.binom(100,100,.05,.2,1,N,"call"); gw = pcf.binom(100,100,.05,.2,1,N,"call",window=True)
    print("binom", N, repr(g.price), "window", repr(gw.price), "kernel s", g.seconds_kernel, gw.seconds_kernel)
# native statistical + timing
BS = 10.450583572185565
for N in (10**7, 10**9):
    g = pcf.mc_eur(100,100,.05,.2,1,N,"call",seed=1)
    print("mc_eur native", N, g.price, "se", g.std_error, "z", (g.price-BS)/g.std_error, "kernel s", g.seconds_kernel, "paths/s %.3e" % (N/g.seconds_kernel))
for N in (10**6, 10**7, 10**8):
    g = pcf.mc_asia(100,100,.05,.2,1,N,252,"call",seed=1)
    print("mc_asia native", N, g.price, "se", g.std_error, "kernel s", g.seconds_kernel, "path-steps/s %.3e" % (N*252/g.seconds_kernel))
for N in (10**6, 10**8):
    g = pcf.mc_eur_multi(100,100,.05,.2,1,N,"call",16,0.5,seed=1)
    print("basket native", N, g.price, "se", g.std_error, "kernel s", g.seconds_kernel, "paths/s %.3e" % (N/g.seconds_kernel))
for N in (10**6, 10**7, 10**8):
    g = pcf.mc_amer(100,100,.05,.2,1,N,50,"put",seed=1)
    print("mc_amer native", N, g.price, "se", g.std_error, "kernel s", g.seconds_kernel, "total s", g.seconds_total, "path-steps/s %.3e" % (N*50/g.seconds_kernel))
print("fp64 peak DFMA/s %.4e" % pcf.fp64_peak(0.3), "hbm copy B/s %.4e" % pcf.hbm_peak(1<<31))
pcf.shutdown()
