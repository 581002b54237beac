"""Generates tests/golden/*.json. Run in the BUILD container only (it needs /root/reference, via the
binaries oracle/Makefile compiles from the unmodified reference sources into oracle/_ref/):

    make -C oracle all && python tests/golden/make_golden.py

basket_vectors.json     outputs of the unmodified src/mc_eur_multi.cpp + include/mvn.h compiled against the stand-in
                        Eigen / Boost.Random headers of oracle/shim/ (oracle/_ref/mc_eur_multi_fn), and the published
                        Serial rows of results/results_mc_eur_multi.csv (statistical pins)
reference_vectors.json  outputs of the compiled reference (function-level harness, %.17g, RNG seed
                        pinned through --wrap=time) + the published rows of
                        /root/reference/results/results_binom_embar.csv
exact_binom.json        exact binomial-formula sums (mpmath, 50 digits) on the reference's own
                        double-precision lattice parameters (u, d, p, q as binom_embar.cpp:19-27
                        derives them), for N beyond what the O(N^2) reference can reach.
Nothing here runs on the GPU box; the tests only read the JSON.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def reference_vectors():
    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {"_how": "oracle/_ref/<prog>_fn <argv>, PCF_FIXED_TIME=<seed>; see tests/golden/make_golden.py",
           "mc_eur": [], "mc_asia": [], "mc_amer": [], "binom_embar": [], "binom_embar_csv": []}
    P1 = (100, 100, 0.05, 0.2, 1)     # BASELINE.json configs
    P2 = (100, 110, 0.02, 0.75, 1)    # the reference Makefile's *_tst parameters (call)
    P3 = (100, 90, 0.02, 0.75, 1)     # ... and put
    for seed, pf, P, N in [(42, "call", P1, 10_000_000), (42, "call", P1, 100_000), (7, "put", P1, 100_001),
                           (42, "call", P2, 50_000), (3, "put", P3, 50_000), (1, "call", P1, 1)]:
        out["mc_eur"].append({"seed": seed, "payoff": pf, "params": P, "N": N,
                              "price": oracle.ref_fn("mc_eur", pf, *P, N, seed=seed)})
    for seed, pf, P, N, M in [(42, "call", P1, 10_000, 252), (42, "call", P1, 100_000, 252), (9, "put", P1, 5_000, 200),
                              (42, "call", P2, 5_000, 200), (5, "put", P3, 4_001, 13), (1, "call", P1, 3, 1)]:
        out["mc_asia"].append({"seed": seed, "payoff": pf, "params": P, "N": N, "M": M,
                               "price": oracle.ref_fn("mc_asia", pf, *P, N, M, seed=seed)})
    for seed, pf, P, N, M in [(42, "put", P1, 100_000, 50), (42, "put", P1, 1_000_000, 50), (42, "call", P2, 100_000, 50),
                              (11, "call", P2, 20_000, 200), (11, "put", P3, 20_000, 200), (2, "call", P1, 10_000, 50),
                              (4, "put", P1, 2, 5), (4, "call", P1, 4, 3), (6, "call", (100, 160, 0.05, 0.2, 1), 2_000, 20)]:
        out["mc_amer"].append({"seed": seed, "payoff": pf, "params": P, "N": N, "M": M,
                               "price": oracle.ref_fn("mc_amer", pf, *P, N, M, seed=seed)})
    for pf, P, N in [("call", P2, 1), ("call", P2, 2), ("call", P2, 3), ("put", P3, 7), ("call", P2, 100), ("call", P2, 800),
                     ("call", P2, 1000), ("put", P3, 1000), ("call", P2, 1001), ("call", P1, 1000), ("put", P1, 999),
                     ("call", P1, 10_000), ("call", P2, 6_400), ("call", P2, 32_000), ("call", P1, 64_000),
                     ("call", P1, 100_000)]:
        out["binom_embar"].append({"payoff": pf, "params": P, "N": N,
                                   "price": oracle.ref_fn("binom_embar", pf, *P, N)})
        print("binom", pf, P, N, out["binom_embar"][-1]["price"], flush=True)
    # published rows: reference results/results_binom_embar.csv, Serial binom_embar rows (10 digits)
    csv = "/root/reference/results/results_binom_embar.csv"
    seen = set()
    for line in open(csv):
        f = line.strip().split(",")
        if len(f) >= 14 and f[0] == "Serial" and f[1] in ("call", "put"):
            key = (f[1], f[7])
            try:
                vals = [float(x) for x in f[2:7]]
                N, price = int(f[7]), float(f[13])
            except ValueError:
                continue
            if key in seen:
                continue
            seen.add(key)
            out["binom_embar_csv"].append({"payoff": f[1], "params": vals, "N": N, "price": price,
                                           "source": "results/results_binom_embar.csv"})
    return out


def exact_binom():
    import mpmath as mp
    mp.mp.dps = 50
    out = {"_how": "mpmath 50-digit sum of C(N,i) p^i q^(N-i) max(cp(S0 u^i d^(N-i) - E),0) e^{-rT} on the "
                   "reference's double u,d,p,q (oracle.binom_params)", "cases": []}
    for pf, (S0, E, r, sigma, T), N in [("call", (100, 100, 0.05, 0.2, 1), 10_000), ("call", (100, 100, 0.05, 0.2, 1), 100_000),
                                        ("put", (100, 100, 0.05, 0.2, 1), 100_000), ("call", (100, 110, 0.02, 0.75, 1), 100_000),
                                        ("call", (100, 100, 0.05, 0.2, 1), 64_000), ("call", (100, 110, 0.02, 0.75, 1), 32_000),
                                        ("call", (100, 100, 0.05, 0.2, 1), 1_000_000), ("put", (100, 90, 0.02, 0.75, 1), 300_001)]:
        u, d, p, q = oracle.binom_params(r, sigma, T, N)
        U, D, Pm, Q = mp.mpf(u), mp.mpf(d), mp.mpf(p), mp.mpf(q)
        lnU, lnD, lnP, lnQ = mp.log(U), mp.log(D), mp.log(Pm), mp.log(Q)
        cp = 1 if pf == "call" else -1
        # log-weights by recurrence in 50-digit arithmetic: w_i = w_{i-1} * (N-i+1)/i * p/q
        total = mp.mpf(0)
        lw = N * lnQ
        for i in range(0, N + 1):
            if i > 0:
                lw += mp.log(mp.mpf(N - i + 1) / i) + lnP - lnQ
            if lw > -900:
                S = S0 * mp.exp(i * lnU + (N - i) * lnD)
                pay = cp * (S - E)
                if pay > 0:
                    total += mp.exp(lw) * pay
        price = mp.exp(-mp.mpf(r) * T) * total
        out["cases"].append({"payoff": pf, "params": [S0, E, r, sigma, T], "N": N, "price": float(price),
                             "price_str": mp.nstr(price, 25)})
        print("exact", pf, N, mp.nstr(price, 25), flush=True)
    return out


def exact_binom_windowed(out):
    """N >= 1e7: only |i - Np| <= 45 sqrt(N) contributes (everything else is below 1e-400)."""
    import math
    import mpmath as mp
    mp.mp.dps = 50
    S0, E, r, sigma, T = 100, 100, 0.05, 0.2, 1
    for N in (10 ** 7, 10 ** 8, 2 ** 31 - 1):
        u, d, p, q = oracle.binom_params(r, sigma, T, N)
        lnU, lnD, lnP, lnQ = (mp.log(mp.mpf(v)) for v in (u, d, p, q))
        W = int(45 * math.sqrt(N)) + 10
        i0, i1 = max(0, int(N * p) - W), min(N, int(N * p) + W)
        for pf, cp in (("call", 1), ("put", -1)):
            lw = mp.loggamma(N + 1) - mp.loggamma(i0 + 1) - mp.loggamma(N - i0 + 1) + i0 * lnP + (N - i0) * lnQ
            tot = mp.mpf(0)
            for i in range(i0, i1 + 1):
                if i > i0:
                    lw += mp.log(mp.mpf(N - i + 1) / i) + lnP - lnQ
                pay = cp * (S0 * mp.exp(i * lnU + (N - i) * lnD) - E)
                if pay > 0:
                    tot += mp.exp(lw) * pay
            price = mp.exp(-mp.mpf(r) * T) * tot
            out["cases"].append({"payoff": pf, "params": [S0, E, r, sigma, T], "N": N, "price": float(price),
                                 "price_str": mp.nstr(price, 25), "windowed": True})
            print("exact(windowed)", pf, N, mp.nstr(price, 25), flush=True)
    out["_how_windowed"] = ("N >= 1e7: same sum restricted to |i - Np| <= 45 sqrt(N) (terms outside are < 1e-400), "
                            "first weight from mpmath loggamma")


def tree_vectors():
    """binom_vanilla_eur / binom_vanilla_amer: outputs of the compiled reference (%.17g) and the
    published Serial_vanilla rows of results/results_binom_embar.csv (10 digits)."""
    out = {"_generated_by": "tests/golden/make_golden.py trees", "binom_vanilla_eur": [], "binom_vanilla_amer": [],
           "binom_vanilla_eur_csv": []}
    cases = [("call", (100, 100, 0.05, 0.2, 1)), ("put", (100, 100, 0.05, 0.2, 1)),
             ("call", (100, 110, 0.02, 0.75, 1)), ("put", (100, 90, 0.02, 0.75, 1)), ("put", (80, 100, 0.1, 0.3, 2.5))]
    for pf, P in cases:
        for N in (1, 2, 3, 7, 31, 32, 33, 64, 100, 191, 192, 193, 1000, 4000, 10000):
            for prog in ("binom_vanilla_eur", "binom_vanilla_amer"):
                if prog.endswith("amer") and N > 4000 and P[0] != 100:
                    continue
                out[prog].append({"payoff": pf, "params": P, "N": N, "price": oracle.ref_fn(prog, pf, *P, N)})
                print(prog, pf, P, N, out[prog][-1]["price"], flush=True)
    for line in open("/root/reference/results/results_binom_embar.csv"):
        f = line.strip().split(",")
        if len(f) >= 14 and f[0] == "Serial_vanilla" and f[1] in ("call", "put"):
            out["binom_vanilla_eur_csv"].append({"payoff": f[1], "params": [float(x) for x in f[2:7]], "N": int(f[7]),
                                                 "price": float(f[13]), "source": "results/results_binom_embar.csv"})
    return out


def basket_vectors():
    """mc_eur_multi: d in {1, 4, 16} as the verdict asks, both payoffs, T != 1 (pins the missing sqrt(T),
    SURVEY F9), a negative rho, and the eigen-decomposition branch of mvn.h:72-76 (rho = 1 at d = 2, where LLT meets a
    zero pivot and the stand-in solver's eigenvalues are exactly {0, 2}). `used_eigen` records the branch the reference
    took (restated by oracle.mvn_transform through the same API calls); a NaN price is recorded as null -- that IS the
    reference's output when an eigenvalue of the semi-definite matrix comes out as -1e-17."""
    import math
    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {"_how": "oracle/_ref/mc_eur_multi_fn <argv>, PCF_FIXED_TIME=<seed>; replay stream = std::mt19937(seed) + "
                   "std::normal_distribution<>{0,1}, Z[n*d + a]; see oracle/shim/ and tests/golden/make_golden.py",
           "mc_eur_multi": [], "mc_eur_multi_csv": []}
    P1 = (100, 100, 0.05, 0.2, 1)
    P4 = (100, 100, 0.1, 0.2, 1)      # the reference Makefile's mc_eur_multi_tst parameters (Makefile:165-169)
    PT = (90, 100, 0.03, 0.4, 2.5)    # T != 1
    for seed, pf, P, N, d, rho in [(42, "call", P1, 100_000, 1, 0.0), (42, "call", P1, 100_000, 4, 0.5),
                                   (42, "call", P1, 50_000, 16, 0.5), (7, "put", P1, 50_001, 16, 0.5),
                                   (42, "call", P4, 1_000_000, 4, 0.5), (3, "put", PT, 20_000, 4, 0.3),
                                   (3, "call", PT, 20_000, 16, 0.9), (5, "call", P1, 30_000, 5, -0.2),
                                   (9, "call", P1, 40_000, 32, 0.5), (1, "call", P1, 1, 3, 0.5),
                                   (42, "call", P1, 100_000, 2, 1.0), (11, "put", PT, 20_000, 2, 1.0),
                                   (42, "call", P1, 10_000, 4, 1.0), (42, "call", P1, 10_000, 3, -0.5)]:
        price = oracle.ref_fn("mc_eur_multi", pf, *P, N, d, rho, seed=seed)
        A, eig = oracle.mvn_transform(d, rho)
        out["mc_eur_multi"].append({"seed": seed, "payoff": pf, "params": P, "N": N, "assets": d, "rho": rho,
                                    "used_eigen": eig, "price": None if math.isnan(price) else price})
        print("mc_eur_multi", pf, P, N, d, rho, eig, price, flush=True)
    for line in open("/root/reference/results/results_mc_eur_multi.csv"):
        f = line.strip().split(",")
        if len(f) >= 14 and f[0] == "Serial" and f[1] in ("call", "put"):
            try:
                out["mc_eur_multi_csv"].append({"payoff": f[1], "params": [float(x) for x in f[2:7]], "N": int(f[7]),
                                                "assets": int(f[10]), "price": float(f[13]),
                                                "source": "results/results_mc_eur_multi.csv"})
            except ValueError:
                continue
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["ref", "exact", "trees", "basket"]
    if "basket" in which:
        json.dump(basket_vectors(), open(os.path.join(HERE, "basket_vectors.json"), "w"), indent=1)
    if "trees" in which:
        json.dump(tree_vectors(), open(os.path.join(HERE, "tree_vectors.json"), "w"), indent=1)
    if "ref" in which:
        json.dump(reference_vectors(), open(os.path.join(HERE, "reference_vectors.json"), "w"), indent=1)
    if "exact" in which:
        ex = exact_binom()
        exact_binom_windowed(ex)
        json.dump(ex, open(os.path.join(HERE, "exact_binom.json"), "w"), indent=1)
