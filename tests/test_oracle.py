"""CPU tests: the oracle (oracle/cpu_ref.cpp) against the reference's golden vectors.

The vectors in tests/golden/reference_vectors.json are outputs of the UNMODIFIED reference compiled
from /root/reference (tests/golden/make_golden.py); the mt19937 normal stream the reference draws is
reproduced by oracle.normals_mt19937 (same libstdc++), so the restatement must match to the last bit
or two (the only licence is the compiler's choice of pow(x,2) vs x*x).
"""
import math

import numpy as np
import pytest

import oracle

BS_CALL = 10.450583572185565  # Black-Scholes call 100/100/.05/.2/1 (reference depr/eur_analytical.R:7-10)
BS_PUT = 5.573526022256971


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert oracle.philox4x32_10(ctr, key) == want


def test_normal_stream_moments():
    z = oracle.normal_stream(123, 1, 0, 20000, 10).reshape(-1)
    assert abs(z.mean()) < 4 / math.sqrt(z.size)
    assert abs(z.std() - 1) < 0.01
    assert abs((z ** 4).mean() - 3) < 0.1
    # pairs from one Philox block are uncorrelated
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 0.01


def test_mc_eur_matches_reference(golden):
    for c in golden["reference_vectors"]["mc_eur"]:
        S0, E, r, sigma, T = c["params"]
        w = oracle.normals_mt19937(c["seed"], math.sqrt(T), c["N"])
        got = oracle.mc_eur(S0, E, r, sigma, T, c["N"], c["payoff"], w)
        assert rel(got, c["price"]) < 1e-15, c


def test_mc_asia_matches_reference(golden):
    for c in golden["reference_vectors"]["mc_asia"]:
        S0, E, r, sigma, T = c["params"]
        w = oracle.normals_mt19937(c["seed"], math.sqrt(T / c["M"]), c["N"] * c["M"])
        got = oracle.mc_asia(S0, E, r, sigma, T, c["N"], c["M"], c["payoff"], w)
        assert rel(got, c["price"]) < 1e-15, c


def test_mc_amer_matches_reference(golden):
    for c in golden["reference_vectors"]["mc_amer"]:
        S0, E, r, sigma, T = c["params"]
        w = oracle.normals_mt19937(c["seed"], math.sqrt(T / c["M"]), c["N"] // 2 * c["M"])
        got = oracle.mc_amer(S0, E, r, sigma, T, c["N"], c["M"], c["payoff"], w)
        assert rel(got, c["price"]) < 1e-15, c


def test_mc_amer_quirk_value():
    # the put prices near 2E - S, not near the true American value 6.09
    N, M = 20000, 50
    w = oracle.normals_mt19937(1, math.sqrt(1 / M), N // 2 * M)
    v = oracle.mc_amer(100, 100, .05, .2, 1, N, M, "put", w)
    assert 90 < v < 94


def test_mc_amer_odd_n_rejected():
    with pytest.raises(ValueError):
        oracle.mc_amer(100, 100, .05, .2, 1, 11, 5, "put", np.zeros(100))


def test_binom_matches_reference(golden):
    for c in golden["reference_vectors"]["binom_embar"]:
        if c["N"] > 10000:
            continue  # O(N^2): covered once at generation time; checked on the GPU side against the fixture
        S0, E, r, sigma, T = c["params"]
        got = oracle.binom(S0, E, r, sigma, T, c["N"], c["payoff"])
        assert rel(got, c["price"]) < 1e-15, c


def test_binom_matches_published_csv(golden):
    # reference results/results_binom_embar.csv, printed with 10 significant digits
    rows = [c for c in golden["reference_vectors"]["binom_embar_csv"] if c["N"] <= 6400]
    assert len(rows) >= 5
    for c in rows:
        S0, E, r, sigma, T = c["params"]
        got = oracle.binom(S0, E, r, sigma, T, c["N"], c["payoff"])
        assert rel(got, c["price"]) < 1e-9, c


def test_trees_match_reference(golden):
    # binom_vanilla_eur / binom_vanilla_amer restated: bit-exact against the compiled reference
    for prog, american in (("binom_vanilla_eur", False), ("binom_vanilla_amer", True)):
        n = 0
        for c in golden["tree_vectors"][prog]:
            if c["N"] > (1000 if american else 4000):
                continue
            assert oracle.binom_tree(*c["params"], c["N"], c["payoff"], american) == c["price"], (prog, c)
            n += 1
        assert n >= 40


def test_trees_match_published_csv(golden):
    rows = [c for c in golden["tree_vectors"]["binom_vanilla_eur_csv"] if c["N"] <= 3200]
    assert len(rows) >= 5
    for c in rows:
        got = oracle.binom_tree(*c["params"], c["N"], c["payoff"], False)
        assert f"{got:.10g}" == f"{c['price']:.10g}", c


def test_binom_converges_to_black_scholes():
    assert abs(oracle.binom(100, 100, .05, .2, 1, 4000, "call") - BS_CALL) < 2e-3
    assert abs(oracle.binom(100, 100, .05, .2, 1, 4000, "put") - BS_PUT) < 2e-3


def test_mc_basket_matches_reference(golden):
    # The unmodified src/mc_eur_multi.cpp + include/mvn.h compiled against the stand-in Eigen/Boost headers (oracle/shim):
    # both branches of mvn.h:68-76, d = 1..32, T != 1, both payoffs; a null price is the reference's NaN.
    n_eig = 0
    for c in golden["basket_vectors"]["mc_eur_multi"]:
        S0, E, r, sigma, T = c["params"]
        d = c["assets"]
        Z = oracle.normals_mt19937(c["seed"], 1.0, c["N"] * d)
        got = oracle.mc_basket(S0, E, r, sigma, T, c["N"], c["payoff"], d, c["rho"], Z)
        assert oracle.mvn_transform(d, c["rho"])[1] == c["used_eigen"], c
        n_eig += c["used_eigen"]
        if c["price"] is None:
            assert math.isnan(got), c
        else:
            assert got == c["price"] or rel(got, c["price"]) < 1e-15, c
    assert n_eig >= 3


def test_mc_basket_published_rows_are_statistically_consistent(golden):
    # results/results_mc_eur_multi.csv (Serial rows, d = 4): each published price lies within 5 standard errors of the
    # restatement on an independent stream of the same N (the reference's seed was time(), so nothing tighter exists)
    rows = [c for c in golden["basket_vectors"]["mc_eur_multi_csv"] if c["N"] <= 1_000_000 and c["assets"] == 4]
    assert len(rows) >= 3
    for c in rows[:6]:
        S0, E, r, sigma, T = c["params"]
        N = c["N"]
        Z = oracle.normals_mt19937(1234 + N, 1.0, N * 4)
        v, s, s2 = oracle.mc_basket(S0, E, r, sigma, T, N, c["payoff"], 4, 0.5, Z, moments=True)   # runscript: rho=0.5
        se = math.exp(-r * T) * math.sqrt(max(s2 / N - (s / N) ** 2, 0) / N)
        assert abs(v - c["price"]) < 5 * math.sqrt(2) * se + 1e-9, (c, v, se)


def test_basket_anchors():
    # analytic anchors next to the compiled-reference vectors above
    N = 200000
    w = oracle.normals_mt19937(5, 1.0, N)
    # d = 1: identical to mc_eur at T = 1
    a = oracle.mc_basket(100, 100, .05, .2, 1, N, "call", 1, 0.3, w)
    b = oracle.mc_eur(100, 100, .05, .2, 1, N, "call", w)
    assert rel(a, b) < 1e-15
    # rho -> 1: every asset follows the first normal => Black-Scholes within MC error
    Z = oracle.normals_mt19937(6, 1.0, N * 4)
    v, s, s2 = oracle.mc_basket(100, 100, .05, .2, 1, N, "call", 4, 1 - 1e-12, Z, moments=True)
    se = math.exp(-.05) * math.sqrt((s2 / N - (s / N) ** 2) / N)
    assert abs(v - BS_CALL) < 4 * se
    # Cholesky factor reproduces the equicorrelation matrix
    L = oracle.chol_equicorr(16, 0.5)
    C = L @ L.T
    assert np.allclose(np.diag(C), 1.0, atol=1e-14) and np.allclose(C - np.diag(np.diag(C)), 0.5 * (1 - np.eye(16)), atol=1e-14)
    with pytest.raises(ValueError):
        oracle.chol_equicorr(4, -0.5)  # not positive definite (rho < -1/(d-1))
    # mvn.h:68-76: LLT fails on the singular matrix -> eigen branch; A A^T still reproduces the matrix
    A, eig = oracle.mvn_transform(2, 1.0)
    assert eig and np.allclose(A @ A.T, np.ones((2, 2)), atol=1e-14)
    A, eig = oracle.mvn_transform(16, 0.5)
    assert not eig and np.array_equal(A, L)


def test_general_basket_reduces_to_reference_basket():
    # with one sigma, one S0, weights 1/d and A = chol(equicorrelation) the general restatement IS the
    # reference loop (src/mc_eur_multi.cpp:23-34) -- same operations up to the order of the 1/d * S0 product
    N, d, rho = 20000, 6, 0.4
    Z = oracle.normals_mt19937(9, 1.0, N * d)
    a = oracle.mc_basket(100, 95, .05, .25, 1, N, "put", d, rho, Z)
    L = oracle.chol_equicorr(d, rho)
    b = oracle.mc_basket_general(100, 95, .05, .25, 1, N, "put", L, np.full(d, 1.0 / d), Z)
    assert rel(b, a) < 1e-13
    # a permutation of (asset, weight, spot, vol, row of A) leaves the price unchanged up to summation order
    rng = np.random.default_rng(3)
    S0 = rng.uniform(80, 120, d); sg = rng.uniform(.1, .4, d); w = rng.dirichlet(np.ones(d))
    B = rng.standard_normal((d, d)); A = np.linalg.cholesky(B @ B.T / d + .2 * np.eye(d))
    perm = rng.permutation(d)
    p0 = oracle.mc_basket_general(S0, 100, .03, sg, 1, N, "call", A, w, Z)
    p1 = oracle.mc_basket_general(S0[perm], 100, .03, sg[perm], 1, N, "call", A[perm], w[perm], Z)
    assert rel(p1, p0) < 1e-12
    # a single asset with weight 1 and A = [[1]] is mc_eur at T = 1
    z = Z[:N]
    assert rel(oracle.mc_basket_general(100, 100, .05, .2, 1, N, "call", np.eye(1), [1.0], z),
               oracle.mc_eur(100, 100, .05, .2, 1, N, "call", z)) < 1e-15


def test_basket_published_statistical_pin():
    # reference results/results_mc_eur_multi.csv: d=4, rho=.5, call 100/100 r=.1 sigma=.2 T=1 -> 11.92 (+-0.01)
    N = 400000
    Z = oracle.normals_mt19937(8, 1.0, N * 4)
    v, s, s2 = oracle.mc_basket(100, 100, .1, .2, 1, N, "call", 4, 0.5, Z, moments=True)
    se = math.exp(-.1) * math.sqrt((s2 / N - (s / N) ** 2) / N)
    assert abs(v - 11.92) < 4 * se + 0.01


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (reference absent)")
def test_compiled_reference_still_agrees():
    # live cross-check when the compiled reference is present (build container, or shipped to the box)
    N, M = 3000, 17
    w = oracle.normals_mt19937(77, math.sqrt(1 / M), N * M)
    assert rel(oracle.mc_asia(100, 100, .05, .2, 1, N, M, "call", w),
               oracle.ref_fn("mc_asia", "call", 100, 100, .05, .2, 1, N, M, seed=77)) < 1e-15
    w = oracle.normals_mt19937(77, math.sqrt(1 / M), N // 2 * M)
    assert rel(oracle.mc_amer(100, 100, .05, .2, 1, N, M, "put", w),
               oracle.ref_fn("mc_amer", "put", 100, 100, .05, .2, 1, N, M, seed=77)) < 1e-15


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (reference absent)")
def test_restatement_equals_compiled_reference_on_random_parameters():
    # Beyond the committed golden vectors: 40 seeded random parameter sets per method, restatement vs the compiled,
    # unmodified reference fed the same mt19937 stream -- bit-equal for the Monte Carlo methods and the trees, and to the
    # reference's own accuracy for the binomial formula (its comb() drifts).
    import numpy as np
    rng = np.random.default_rng(20240229)
    for k in range(40):
        S0 = float(rng.uniform(50, 150)); E = float(rng.uniform(50, 150))
        r = float(rng.uniform(0.0, 0.1)); sigma = float(rng.uniform(0.05, 0.8)); T = float(rng.uniform(0.25, 3.0))
        pf = "call" if k % 2 else "put"
        seed = int(rng.integers(1, 10**6))
        N = int(rng.integers(2, 300)) * 2
        M = int(rng.integers(2, 40))
        w = oracle.normals_mt19937(seed, math.sqrt(T), N)
        assert oracle.mc_eur(S0, E, r, sigma, T, N, pf, w) == oracle.ref_fn("mc_eur", pf, S0, E, r, sigma, T, N, seed=seed)
        w = oracle.normals_mt19937(seed, math.sqrt(T / M), N * M)
        assert oracle.mc_asia(S0, E, r, sigma, T, N, M, pf, w) == \
            oracle.ref_fn("mc_asia", pf, S0, E, r, sigma, T, N, M, seed=seed)
        w = oracle.normals_mt19937(seed, math.sqrt(T / M), N // 2 * M)
        try:
            want = oracle.ref_fn("mc_amer", pf, S0, E, r, sigma, T, N, M, seed=seed)
        except Exception:           # the reference aborts on a singular regression (det <= 0): so must the restatement
            with pytest.raises(Exception):
                oracle.mc_amer(S0, E, r, sigma, T, N, M, pf, w)
        else:
            assert oracle.mc_amer(S0, E, r, sigma, T, N, M, pf, w) == want
        d = int(rng.integers(1, 20))
        rho = float(rng.uniform(-1.0 / max(d - 1, 1) + 1e-3, 1.0)) if k % 5 else 1.0   # every 5th: the eigen branch
        Z = oracle.normals_mt19937(seed, 1.0, N * d)
        got, want = oracle.mc_basket(S0, E, r, sigma, T, N, pf, d, rho, Z), \
            oracle.ref_fn("mc_eur_multi", pf, S0, E, r, sigma, T, N, d, rho, seed=seed)
        assert got == want or (math.isnan(got) and math.isnan(want)), (k, d, rho, got, want)
        Nt = int(rng.integers(1, 400))
        assert oracle.binom_tree(S0, E, r, sigma, T, Nt, pf, False) == \
            oracle.ref_fn("binom_vanilla_eur", pf, S0, E, r, sigma, T, Nt)
        assert oracle.binom_tree(S0, E, r, sigma, T, Nt, pf, True) == \
            oracle.ref_fn("binom_vanilla_amer", pf, S0, E, r, sigma, T, Nt)
        want = oracle.ref_fn("binom_embar", pf, S0, E, r, sigma, T, Nt)
        assert rel(oracle.binom(S0, E, r, sigma, T, Nt, pf), want) < 1e-11 or abs(want) < 1e-9
