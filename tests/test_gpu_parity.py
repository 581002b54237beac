"""GPU parity tests (-m gpu): libpcf.so through its C ABI against the oracle and the golden vectors.

Tolerances (BASELINE.json north_star):
  replay mode  -- same normal stream as the reference: 1e-12 relative (only summation order and
                  last-ulp libm differences remain);
  native mode  -- within 3 standard errors of the reference / Black-Scholes;
  binomial sum -- 1e-10 relative against binom_embar, and 1e-12 against the exact sum.
"""
import math

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

REPLAY_TOL = 1e-12
BINOM_TOL = 1e-10
BS_CALL = 10.450583572185565
BS_PUT = 5.573526022256971
P1 = (100, 100, 0.05, 0.2, 1)


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


# ---------------------------------------------------------------------------------------------------
# generator
def test_philox_known_answers_on_device(gpu):
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert gpu.philox4x32_10(ctr, key) == want


def test_normal_stream_matches_cpu_restatement(gpu):
    for seed, stream, i0, cnt, T in [(7, 1, 1000, 4096, 5), (2 ** 40 + 3, 3, 2 ** 33, 1000, 50), (0, 0, 0, 1, 1)]:
        g = gpu.normal_stream(seed, stream, i0, cnt, T, 0.5)
        o = oracle.normal_stream(seed, stream, i0, cnt, T, 0.5)
        assert np.abs(g - o).max() < 5e-15


def test_normal_stream_distribution(gpu):
    z = gpu.normal_stream(99, 1, 0, 200000, 20).reshape(-1)
    n = z.size
    assert abs(z.mean()) < 4 / math.sqrt(n)
    assert abs(z.var() - 1) < 4 * math.sqrt(2 / n)
    assert abs((z ** 3).mean()) < 4 * math.sqrt(15 / n)
    assert abs((z ** 4).mean() - 3) < 4 * math.sqrt(96 / n)
    # tail mass beyond 3 sigma
    p3 = 2 * 0.0013498980316301
    assert abs((np.abs(z) > 3).mean() - p3) < 4 * math.sqrt(p3 / n)


# ---------------------------------------------------------------------------------------------------
# replay parity against the compiled reference's outputs (golden) and against the oracle
def test_mc_eur_replay_golden(gpu, golden):
    for c in golden["reference_vectors"]["mc_eur"]:
        S0, E, r, sigma, T = c["params"]
        w = oracle.normals_mt19937(c["seed"], math.sqrt(T), c["N"])
        g = gpu.mc_eur(S0, E, r, sigma, T, c["N"], c["payoff"], replay=w)
        assert rel(g.price, c["price"]) < REPLAY_TOL, c
        assert g.n == c["N"] and g.units == c["N"] and g.launches >= 1


def test_mc_asia_replay_golden(gpu, golden):
    for c in golden["reference_vectors"]["mc_asia"]:
        S0, E, r, sigma, T = c["params"]
        w = oracle.normals_mt19937(c["seed"], math.sqrt(T / c["M"]), c["N"] * c["M"])
        g = gpu.mc_asia(S0, E, r, sigma, T, c["N"], c["M"], c["payoff"], replay=w)
        assert rel(g.price, c["price"]) < REPLAY_TOL, c
        assert g.units == c["N"] * c["M"]


def test_mc_amer_replay_golden(gpu, golden):
    for c in golden["reference_vectors"]["mc_amer"]:
        S0, E, r, sigma, T = c["params"]
        w = oracle.normals_mt19937(c["seed"], math.sqrt(T / c["M"]), c["N"] // 2 * c["M"])
        g = gpu.mc_amer(S0, E, r, sigma, T, c["N"], c["M"], c["payoff"], replay=w)
        assert rel(g.price, c["price"]) < REPLAY_TOL, c
        # paths [+ pad] + state fill + one fused sweep kernel per date (the kernel of date 1 also forms the final sum)
        assert g.launches == 2 + c["M"] + (1 if c["N"] % 2944 else 0)  # paths, date fill, one sweep per date (+ padding to whole tiles)


def test_mc_amer_few_itm_branches(gpu):
    # deep out-of-the-money call: most dates have 0, 1 or 2 paths in the money (mc_amer.cpp:73-83)
    for seed, N, M, E in [(6, 2000, 20, 160), (8, 500, 30, 150), (9, 64, 40, 135)]:
        w = oracle.normals_mt19937(seed, math.sqrt(1 / M), N // 2 * M)
        o = oracle.mc_amer(100, E, .05, .2, 1, N, M, "call", w)
        g = gpu.mc_amer(100, E, .05, .2, 1, N, M, "call", replay=w)
        assert rel(g.price, o) < REPLAY_TOL or abs(g.price - o) < 1e-15


def test_mc_amer_date_width_boundary(gpu):
    # exercise dates are stored in 1 byte up to M = 127 and in 2 bytes beyond: both sides of the switch, both rules,
    # paths that exercise early, late and never (every one of them re-gathers paths[when][n], mc_amer.cpp:50)
    for M in (126, 127, 128, 129, 300):
        for pf, P in (("put", P1), ("call", (100, 110, 0.02, 0.75, 1))):
            N = 3002
            w = oracle.normals_mt19937(100 + M, math.sqrt(P[4] / M), N // 2 * M)
            for lsm in (False, True):
                o = oracle.mc_amer(*P, N, M, pf, w, lsm=lsm)
                g = gpu.mc_amer(*P, N, M, pf, replay=w, lsm=lsm)
                assert rel(g.price, o) < REPLAY_TOL, (M, pf, lsm)


def test_mc_amer_textbook_lsm_mode(gpu):
    # PCF_FLAG_AMER_LSM -- true payoff against the fitted continuation value. Replay parity with
    # the oracle's restatement of that rule, and the price lands next to the binomial tree's American put
    # (reference bin/binom_vanilla_amer put 100 100 0.05 0.2 1 2000 -> 6.090232044; 50 exercise dates and a
    # quadratic basis sit ~0.5 % below it), not at the reference rule's 91.9.
    for pf, P, N, M, seed in [("put", P1, 100_000, 50, 3), ("call", (100, 110, 0.02, 0.75, 1), 50_000, 20, 4),
                              ("put", (100, 90, 0.02, 0.75, 1), 20_002, 200, 5)]:
        w = oracle.normals_mt19937(seed, math.sqrt(P[4] / M), N // 2 * M)
        o = oracle.mc_amer(*P, N, M, pf, w, lsm=True)
        g = gpu.mc_amer(*P, N, M, pf, replay=w, lsm=True)
        assert rel(g.price, o) < REPLAY_TOL, (pf, N, M)
    g = gpu.mc_amer(*P1, 4_000_000, 50, "put", seed=9, lsm=True)
    assert 6.00 < g.price < 6.0903 and g.std_error < 0.01
    d = gpu.mc_amer(*P1, 4_000_000, 50, "put", seed=9)
    assert d.price > 90  # the default stays the reference's rule


def test_basket_replay_vs_oracle(gpu):
    for d, rho, N, pf in [(16, 0.5, 50000, "call"), (4, 0.5, 30000, "call"), (3, 0.2, 10001, "put"), (1, 0.0, 5000, "call"),
                          (32, 0.9, 4000, "put"), (7, -0.1, 3000, "call")]:
        Z = oracle.normals_mt19937(42 + d, 1.0, N * d)
        o = oracle.mc_basket(100, 100, .05, .2, 1, N, pf, d, rho, Z)
        g = gpu.mc_eur_multi(100, 100, .05, .2, 1, N, pf, d, rho, replay=Z)
        assert rel(g.price, o) < REPLAY_TOL, (d, rho)


def test_basket_replay_vs_compiled_reference(gpu, golden):
    # tests/golden/basket_vectors.json: the UNMODIFIED src/mc_eur_multi.cpp + include/mvn.h (compiled against oracle/shim)
    # run on std::mt19937(seed) normals; the GPU consumes the same stream, Z[n*d + a]. Covers both branches of
    # mvn.h:68-76 through the reference's own entry point (pcf_mc_eur_multi), d = 1..32, T != 1.
    n_eig = 0
    for c in golden["basket_vectors"]["mc_eur_multi"]:
        S0, E, r, sigma, T = c["params"]
        d, N = c["assets"], c["N"]
        Z = oracle.normals_mt19937(c["seed"], 1.0, N * d)
        g = gpu.mc_eur_multi(S0, E, r, sigma, T, N, c["payoff"], d, c["rho"], replay=Z)
        if c["price"] is not None:
            assert abs(g.price - c["price"]) <= REPLAY_TOL * max(abs(c["price"]), 1e-3), c
            n_eig += c["used_eigen"]
        else:
            # the reference printed NaN: an eigenvalue of the semi-definite matrix came out as -1e-17 and mvn.h:75 took
            # its square root. The product counts it as 0 (include/pcf.h) and must agree with the restatement fed the
            # product's own factor.
            assert c["used_eigen"] and math.isfinite(g.price)
            cov = np.full((d, d), c["rho"]); np.fill_diagonal(cov, 1.0)
            A, eig = gpu.normal_transform(cov)
            assert eig and np.allclose(A @ A.T, cov, atol=1e-12)
            o = oracle.mc_basket_general(S0, E, r, sigma, T, N, c["payoff"], A, np.full(d, 1.0 / d), Z)
            assert rel(g.price, o) < REPLAY_TOL, c
    assert n_eig >= 2
    # rho = 1 (every asset driven by the same combination of normals) collapses to one asset: Black-Scholes, native mode
    g = gpu.mc_eur_multi(*P1, 4_000_000, "call", 8, 1.0, seed=12)
    assert abs(g.price - BS_CALL) < 3 * g.std_error
    # rho = -1/(d-1): the basket's driving noise sums to zero variance in the equal-weight direction; finite, below BS
    g = gpu.mc_eur_multi(*P1, 1_000_000, "call", 3, -0.5, seed=12)
    assert math.isfinite(g.price) and 0 < g.price < BS_CALL


def test_basket_equicorrelation_fast_path_is_bit_identical(gpu):
    # the constant-column shortcut performs the same chain of FMAs as the general triangular product
    # (PCF_FLAG_BASKET_GENERAL sends the reference's basket through the general kernel)
    for d, rho, N in [(16, 0.5, 300_001), (5, -0.1, 100_000), (32, 0.3, 50_000), (1, 0.0, 10_000), (2, 0.9, 10_001)]:
        fast = gpu.mc_eur_multi(*P1, N, "call", d, rho, seed=77)
        gen = gpu.mc_eur_multi(*P1, N, "call", d, rho, seed=77, general=True)
        assert rel(fast.sum, gen.sum) < 1e-14 and rel(fast.sumsq, gen.sumsq) < 1e-14, (d, rho)


def test_general_basket_replay_vs_oracle(gpu):
    # per-asset spots/vols/weights and a general covariance; oracle and GPU share the factor A
    rng = np.random.default_rng(11)
    for d, N, pf, kind in [(16, 40000, "call", "chol"), (5, 20001, "put", "chol"), (32, 3000, "call", "chol"),
                           (9, 20000, "call", "semidef"), (4, 10000, "put", "rho1"), (1, 5000, "call", "chol")]:
        S0 = rng.uniform(80, 120, d); sg = rng.uniform(.1, .4, d); w = rng.dirichlet(np.ones(d))
        if kind == "chol":
            B = rng.standard_normal((d, d)); cov = B @ B.T / d + .2 * np.eye(d)
        elif kind == "semidef":
            B = rng.standard_normal((d, 3)); cov = B @ B.T  # rank 3: Cholesky fails, mvn.h:72-76 fallback
        else:
            cov = np.ones((d, d))
        A, eig = gpu.normal_transform(cov)
        assert eig == (kind != "chol") and np.allclose(A @ A.T, cov, atol=1e-12)
        Z = oracle.normals_mt19937(100 + d, 1.0, N * d)
        o, so, so2 = oracle.mc_basket_general(S0, 100, .03, sg, 1, N, pf, A, w, Z, moments=True)
        g = gpu.mc_basket(S0, 100, .03, sg, 1, N, pf, d, weights=w, cov=cov, replay=Z)
        assert rel(g.price, o) < REPLAY_TOL and rel(g.sumsq, so2) < REPLAY_TOL, (d, kind)
        g2 = gpu.mc_basket(S0, 100, .03, sg, 1, N, pf, d, weights=w, transform=A, replay=Z)
        assert g2.sum == g.sum
    # every array NULL = the reference's basket: same numbers as pcf_mc_eur_multi on the same Philox stream
    a = gpu.mc_basket(100, 100, .05, .2, 1, 200_001, "call", 16, rho=0.5, seed=9)
    b = gpu.mc_eur_multi(100, 100, .05, .2, 1, 200_001, "call", 16, 0.5, seed=9)
    assert rel(a.sum, b.sum) < 1e-13 and rel(a.sumsq, b.sumsq) < 1e-13
    # native mode, heterogeneous basket: GPU price within 3 standard errors of the oracle on an independent stream
    d, N = 8, 400_000
    S0 = rng.uniform(80, 120, d); sg = rng.uniform(.1, .4, d); w = rng.dirichlet(np.ones(d))
    B = rng.standard_normal((d, d)); cov = B @ B.T / d + .2 * np.eye(d)
    A, _ = gpu.normal_transform(cov)
    g = gpu.mc_basket(S0, 100, .03, sg, 1, N, "call", d, weights=w, cov=cov, seed=4)
    o, so, so2 = oracle.mc_basket_general(S0, 100, .03, sg, 1, N, "call", A, w,
                                          oracle.normals_mt19937(77, 1.0, N * d), moments=True)
    se_o = math.exp(-.03) * math.sqrt((so2 / N - (so / N) ** 2) / N)
    assert abs(g.price - o) < 3 * math.hypot(g.std_error, se_o)
    with pytest.raises(ValueError):
        gpu.mc_basket(100, 100, .05, .2, 1, 1000, "call", 2, cov=np.array([[1.0, 2.0], [2.0, 1.0]]))


def test_basket_cholesky_matches_oracle(gpu):
    for d, rho in [(16, 0.5), (32, 0.99), (5, -0.2)]:
        assert np.abs(gpu.chol_equicorr(d, rho) - oracle.chol_equicorr(d, rho)).max() < 1e-15


# ---------------------------------------------------------------------------------------------------
# native mode: the oracle fed the kernels' own Philox normals must reproduce the GPU price
def test_native_equals_oracle_on_dumped_stream(gpu):
    seed = 5
    N = 100001
    z = gpu.normal_stream(seed, gpu.STREAM_EUR, 0, (N + 1) // 2, 2).reshape(-1)[:N]
    assert rel(gpu.mc_eur(*P1, N, "put", seed=seed).price, oracle.mc_eur(*P1, N, "put", z)) < REPLAY_TOL
    N, M = 5000, 253  # odd M: last Philox block half used
    z = gpu.normal_stream(seed, gpu.STREAM_ASIA, 0, N, M, math.sqrt(1 / M))
    assert rel(gpu.mc_asia(*P1, N, M, "call", seed=seed).price, oracle.mc_asia(*P1, N, M, "call", z)) < REPLAY_TOL
    N, d = 20000, 16
    z = gpu.normal_stream(seed, gpu.STREAM_BASKET, 0, N, d)
    assert rel(gpu.mc_eur_multi(*P1, N, "call", d, 0.5, seed=seed).price,
               oracle.mc_basket(*P1, N, "call", d, 0.5, z)) < REPLAY_TOL
    N, M = 20000, 50
    z = gpu.normal_stream(seed, gpu.STREAM_AMER, 0, N // 2, M, math.sqrt(1 / M))
    assert rel(gpu.mc_amer(*P1, N, M, "put", seed=seed).price, oracle.mc_amer(*P1, N, M, "put", z)) < REPLAY_TOL
    N, M = 6000, 7
    z = gpu.normal_stream(seed, gpu.STREAM_AMER, 0, N // 2, M, math.sqrt(1 / M))
    assert rel(gpu.mc_amer(100, 110, .02, .75, 1, N, M, "call", seed=seed).price,
               oracle.mc_amer(100, 110, .02, .75, 1, N, M, "call", z)) < REPLAY_TOL


def test_native_within_three_standard_errors(gpu, golden):
    g = gpu.mc_eur(*P1, 10_000_000, "call", seed=20240229)   # BASELINE config 1
    assert abs(g.price - BS_CALL) < 3 * g.std_error and 0.003 < g.std_error < 0.006
    g = gpu.mc_eur(*P1, 10_000_000, "put", seed=1)
    assert abs(g.price - BS_PUT) < 3 * g.std_error
    # Asian: against the reference's own run (golden, N=1e5 paths => its error bar dominates)
    ref = [c for c in golden["reference_vectors"]["mc_asia"] if c["N"] == 100_000 and c["M"] == 252][0]
    g = gpu.mc_asia(*P1, 4_000_000, 252, "call", seed=3)
    se_ref = g.std_error * math.sqrt(4_000_000 / ref["N"])
    assert abs(g.price - ref["price"]) < 3 * math.hypot(g.std_error, se_ref)
    # American (reference's scheme): against the reference's own 1e6-path run
    ref = [c for c in golden["reference_vectors"]["mc_amer"] if c["N"] == 1_000_000][0]
    g = gpu.mc_amer(*P1, 4_000_000, 50, "put", seed=3)
    se_ref = g.std_error * math.sqrt(4_000_000 / ref["N"])
    assert abs(g.price - ref["price"]) < 3 * math.hypot(g.std_error, se_ref)
    # basket: d = 1 is the European call; rho -> 1 collapses to Black-Scholes
    g = gpu.mc_eur_multi(*P1, 4_000_000, "call", 1, 0.0, seed=4)
    assert abs(g.price - BS_CALL) < 3 * g.std_error
    g = gpu.mc_eur_multi(*P1, 2_000_000, "call", 16, 1 - 1e-12, seed=4)
    assert abs(g.price - BS_CALL) < 3 * g.std_error
    # published statistical pin, reference results/results_mc_eur_multi.csv (d=4, rho=.5, r=.1): 11.92
    g = gpu.mc_eur_multi(100, 100, .1, .2, 1, 20_000_000, "call", 4, 0.5, seed=4)
    assert abs(g.price - 11.92) < 3 * g.std_error + 0.01


# ---------------------------------------------------------------------------------------------------
# binomial term sum
def test_binom_matches_reference_vectors(gpu, golden):
    exact = {(c["payoff"], tuple(c["params"]), c["N"]): c["price"] for c in golden["exact_binom"]["cases"]}
    for c in golden["reference_vectors"]["binom_embar"]:
        S0, E, r, sigma, T = c["params"]
        g = gpu.binom(S0, E, r, sigma, T, c["N"], c["payoff"])
        assert g.units == c["N"] + 1
        if c["N"] <= 32000:
            assert rel(g.price, c["price"]) < BINOM_TOL, c
        else:
            # comb() re-sums up to N logarithms per term and the reference drifts away from the exact sum of
            # its own lattice: 5.0e-10 at N = 64000, 1.2e-10 at N = 1e5 (tests/golden/
            # exact_binom.json). Beyond 32000 steps the gate is the exact sum, and the distance to the
            # reference must be the reference's own error.
            ex = exact[(c["payoff"], tuple(float(x) for x in c["params"]), c["N"])]
            assert rel(g.price, ex) < 1e-12, c
            assert abs(rel(g.price, c["price"]) - rel(c["price"], ex)) < 1e-12, c


def test_binom_matches_published_csv(gpu, golden):
    for c in golden["reference_vectors"]["binom_embar_csv"]:
        S0, E, r, sigma, T = c["params"]
        g = gpu.binom(S0, E, r, sigma, T, c["N"], c["payoff"])
        # printed with 10 significant digits; beyond 32000 steps the published value itself is off the exact
        # sum of its lattice by up to 5.5e-9 (N = 80000: published 26.61220685, exact 26.6122069974621652,
        # mpmath) because comb() accumulates rounding error
        assert rel(g.price, c["price"]) < (2e-9 if c["N"] <= 32000 else 1e-8), c
    g = gpu.binom(100, 110, 0.02, 0.75, 1, 80000, "call")
    assert rel(g.price, 26.612206997462165204) < 1e-12


def test_binom_matches_exact_sum(gpu, golden):
    for c in golden["exact_binom"]["cases"]:
        S0, E, r, sigma, T = c["params"]
        g = gpu.binom(S0, E, r, sigma, T, c["N"], c["payoff"])
        assert rel(g.price, c["price"]) < 1e-12, c
        gw = gpu.binom(S0, E, r, sigma, T, c["N"], c["payoff"], window=True)
        assert rel(gw.price, g.price) < 1e-14


def test_binom_screening_pass_is_bit_identical(gpu):
    # the two-logarithm screen (binom_math.cuh pair_dead) only settles pairs the full routine evaluates to exactly 0.0
    for P, pf in ((P1, "call"), (P1, "put"), ((100, 110, 0.02, 0.75, 1), "call"), ((50, 60, 0.1, 0.9, 2.5), "put")):
        for N in (1, 2, 3, 10, 101, 1000, 4097, 100_000, 1_000_003, 30_000_000):
            a = gpu.binom(*P, N, pf)
            b = gpu.binom(*P, N, pf, screen=False)
            assert a.sum == b.sum and a.units == b.units == N + 1, (P, pf, N, a.sum, b.sum)


def test_binom_full_size_properties(gpu, golden):
    # BASELINE config 2 sizes (N = 1e5 .. 1e8, and the int limit): no NaN where the reference overflows
    # (SURVEY F3); exact agreement with the 50-digit sum on the reference's own lattice doubles; put-call
    # parity of that lattice. Convergence to Black-Scholes is only 1/N up to N ~ 1e6: beyond that the
    # reference's u, d, p (sqrt(beta^2 - 1)) carry ~1e-7 of cancellation error and the LATTICE ITSELF
    # drifts (exact sum at N = 2^31-1: 10.4505921, Black-Scholes 10.4505836) -- reproducing that drift is parity.
    exact = {(c["payoff"], c["N"]): c["price"] for c in golden["exact_binom"]["cases"] if tuple(c["params"]) == P1}
    for N in (10 ** 5, 10 ** 6, 10 ** 7, 10 ** 8, 2 ** 31 - 1):
        c = gpu.binom(*P1, N, "call").price
        p = gpu.binom(*P1, N, "put").price
        assert math.isfinite(c) and math.isfinite(p)
        assert rel(c, exact[("call", N)]) < 1e-12, (N, c)
        if ("put", N) in exact:
            assert rel(p, exact[("put", N)]) < 1e-12, (N, p)
        assert abs(c - BS_CALL) < (2.0 / N + 1e-9 if N <= 10 ** 6 else 1e-5)
        assert abs((c - p) - (100 - 100 * math.exp(-0.05))) < (1e-9 if N <= 10 ** 6 else 3e-5)
        w = gpu.binom(*P1, N, "call", window=True).price
        assert rel(w, c) < 1e-14


# ---------------------------------------------------------------------------------------------------
# size-independent properties at larger sizes, edge cases, error behaviour
def test_trees_match_reference_vectors(gpu, golden):
    # SURVEY 8f.1. Every node is formed with the reference's operations in the reference's order; the only liberty is
    # the division by the constant R (csrc/tree_kernels.cu), exact unless a quotient lies within 2^-105 of a rounding
    # boundary. Stated tolerance 1e-13 relative; observed: bit-equal on every vector.
    exact = total = 0
    for prog, fn in (("binom_vanilla_eur", gpu.binom_vanilla_eur), ("binom_vanilla_amer", gpu.binom_vanilla_amer)):
        for c in golden["tree_vectors"][prog]:
            g = fn(*c["params"], c["N"], c["payoff"])
            assert abs(g.price - c["price"]) <= 1e-13 * max(1.0, abs(c["price"])), (prog, c, g.price)
            assert g.units == c["N"] * (c["N"] + 1) // 2
            exact += g.price == c["price"]
            total += 1
    assert exact == total, (exact, total)


def test_trees_match_published_csv(gpu, golden):
    # reference results/results_binom_embar.csv, Serial_vanilla rows, N = 100 .. 100000 (45.6 s there at N = 1e5)
    for c in golden["tree_vectors"]["binom_vanilla_eur_csv"]:
        g = gpu.binom_vanilla_eur(*c["params"], c["N"], c["payoff"])
        assert f"{g.price:.10g}" == f"{c['price']:.10g}", c


def test_trees_launch_shape_independence_and_properties(gpu):
    # the tiling must not change a single bit (same operations per node): the CTA-cooperative kernel with the shape picked
    # per launch (default) against the warp-trapezoid kernel (PCF_FLAG_TREE_WARP). A `make lib TUNING=1` build pins
    # every other shape through PCF_TREE (tools/tune_tree4.py asserts the same equality over all of them).
    P = (100, 100, .05, .2, 1)
    for N in (1, 5, 31, 97, 1000, 5003, 20011, 70001):
        for pf in ("call", "put"):
            e = gpu.binom_vanilla_eur(*P, N, pf).price
            a = gpu.binom_vanilla_amer(*P, N, pf).price
            assert (e, a) == (gpu.binom_vanilla_eur(*P, N, pf, warp_tiling=True).price,
                              gpu.binom_vanilla_amer(*P, N, pf, warp_tiling=True).price), (N, pf)
    # small trees against the oracle, including N not a multiple of anything
    for N in (1, 2, 3, 17, 200, 777):
        for pf in ("call", "put"):
            assert gpu.binom_vanilla_eur(*P, N, pf).price == oracle.binom_tree(*P, N, pf, False)
            assert gpu.binom_vanilla_amer(*P, N, pf).price == oracle.binom_tree(*P, N, pf, True)
    # lattices whose rounded probabilities leave [0, 1] (sigma -> 0): the CTA kernel's single max would differ from the
    # reference's max(continuation, max(payoff, 0)), so the host routes them to the warp kernel -- still bit-equal
    for r_, sig_, N in ((0.1, 1e-7, 1000), (-0.05, 1e-7, 100)):
        for pf in ("call", "put"):
            for E_ in (100, 95, 105):
                Pq = (100, E_, r_, sig_, 1)
                assert gpu.binom_vanilla_amer(*Pq, N, pf).price == oracle.binom_tree(*Pq, N, pf, True), (Pq, N, pf)
                assert gpu.binom_vanilla_eur(*Pq, N, pf).price == oracle.binom_tree(*Pq, N, pf, False), (Pq, N, pf)
    # wide trees walk every rule of the per-launch shape table (csrc/tree_kernels.cu: tree_pick_shape)
    N = 600_000
    auto = (gpu.binom_vanilla_eur(*P, N, "put").price, gpu.binom_vanilla_amer(*P, N, "put").price)
    assert auto == (gpu.binom_vanilla_eur(*P, N, "put", warp_tiling=True).price,
                    gpu.binom_vanilla_amer(*P, N, "put", warp_tiling=True).price)
    # full size (the reference needs 45 s per tree here): the European tree and the binomial formula price the same
    # lattice; early exercise is worth something for the put and nothing for the call (r > 0, no dividends)
    N = 100_000
    e_put, a_put = gpu.binom_vanilla_eur(*P, N, "put").price, gpu.binom_vanilla_amer(*P, N, "put").price
    e_call, a_call = gpu.binom_vanilla_eur(*P, N, "call").price, gpu.binom_vanilla_amer(*P, N, "call").price
    assert rel(e_call, gpu.binom(*P, N, "call").price) < 1e-9 and rel(e_put, gpu.binom(*P, N, "put").price) < 1e-9
    assert 6.08 < a_put < 6.10 and a_put > e_put + 0.4          # binom_vanilla_amer put 100/100/.05/.2/1 -> 6.0903
    assert rel(a_call, e_call) < 1e-12
    assert rel(e_call - e_put, 100 - 100 * math.exp(-0.05)) < 1e-8   # put-call parity on the lattice


def test_determinism_and_homogeneity(gpu):
    a = gpu.mc_asia(*P1, 2_000_000, 252, "call", seed=11)
    b = gpu.mc_asia(*P1, 2_000_000, 252, "call", seed=11)
    assert a.price == b.price and a.sumsq == b.sumsq          # fixed reduction order
    c = gpu.mc_asia(200, 200, 0.05, 0.2, 1, 2_000_000, 252, "call", seed=11)
    assert rel(c.price, 2 * a.price) < 1e-12                  # payoff is homogeneous of degree 1 in (S0, E)
    d = gpu.mc_asia(*P1, 2_000_000, 252, "call", seed=12)
    assert d.price != a.price
    e1 = gpu.mc_amer(*P1, 1_000_000, 50, "put", seed=11)
    e2 = gpu.mc_amer(*P1, 1_000_000, 50, "put", seed=11)
    assert e1.price == e2.price


def test_full_size_properties(gpu, golden):
    """BASELINE.json's sizes (configs 3, 4, 5 and the 2e9-path European run of the roofline figure), where the oracle
    cannot follow: size-independent properties instead -- exact homogeneity, run-to-run bit equality, closed forms
    within 3 standard errors, agreement with the smaller runs that ARE pinned against the reference."""
    # config 3: 1e9 paths x 252 dates. Scaling (S0, E) by 2 is exact in binary FP: every path doubles bit for bit.
    a = gpu.mc_asia(*P1, 1_000_000_000, 252, "call", seed=20240229)
    b = gpu.mc_asia(200, 200, 0.05, 0.2, 1, 1_000_000_000, 252, "call", seed=20240229)
    assert a.n == 10 ** 9 and a.units == 252 * 10 ** 9
    assert b.sum == 2 * a.sum and b.sumsq == 4 * a.sumsq
    small = gpu.mc_asia(*P1, 4_000_000, 252, "call", seed=3)      # pinned against the reference's run above
    assert abs(a.price - small.price) < 3 * math.hypot(a.std_error, small.std_error)
    assert a.std_error < small.std_error / 15                       # 250x the paths: error bar shrinks ~15.8x
    # the European roofline run: 2e9 paths against Black-Scholes (standard error ~3.3e-4)
    e = gpu.mc_eur(*P1, 2_000_000_000, "call", seed=20240229)
    assert abs(e.price - BS_CALL) < 3 * e.std_error and e.std_error < 4e-4
    # config 4: d = 16, 1e9 paths; rho -> 1 collapses the basket to one asset => Black-Scholes
    g = gpu.mc_eur_multi(*P1, 1_000_000_000, "call", 16, 1 - 1e-12, seed=4)
    assert abs(g.price - BS_CALL) < 3 * g.std_error and g.std_error < 6e-4
    g1 = gpu.mc_eur_multi(*P1, 1_000_000_000, "call", 16, 0.5, seed=20240229)
    g2 = gpu.mc_eur_multi(*P1, 20_000_000, "call", 16, 0.5, seed=5)
    assert abs(g1.price - g2.price) < 3 * math.hypot(g1.std_error, g2.std_error)
    assert g1.price < BS_CALL                                       # diversification lowers the basket's volatility
    # config 5: 1e8 paths x 50 dates (40 GB path store): bit-reproducible, floored at immediate exercise, and within
    # the error bars of the 4e6-path run that is pinned against the reference's own 1e6-path result
    x = gpu.mc_amer(*P1, 100_000_000, 50, "put", seed=20240229)
    y = gpu.mc_amer(*P1, 100_000_000, 50, "put", seed=20240229)
    assert x.price == y.price and x.sumsq == y.sumsq and x.units == 50 * 10 ** 8
    z = gpu.mc_amer(*P1, 4_000_000, 50, "put", seed=3)
    assert abs(x.price - z.price) < 3 * math.hypot(x.std_error, z.std_error)
    ref = [c for c in golden["reference_vectors"]["mc_amer"] if c["N"] == 1_000_000][0]
    assert abs(x.price - ref["price"]) < 3 * z.std_error * 2        # reference run: 1e6 paths, error bar 2x z's


def test_put_call_parity_same_stream(gpu):
    N = 5_000_000
    c = gpu.mc_eur(*P1, N, "call", seed=21)
    p = gpu.mc_eur(*P1, N, "put", seed=21)
    f = gpu.mc_eur(100, 0.0, 0.05, 0.2, 1, N, "call", seed=21)   # discounted mean of S_T
    assert rel(c.price - p.price, f.price - 100 * math.exp(-0.05)) < 1e-10


def test_edge_cases(gpu):
    # smallest inputs
    w = oracle.normals_mt19937(1, 1.0, 1)
    assert rel(gpu.mc_eur(*P1, 1, "call", replay=w).price, oracle.mc_eur(*P1, 1, "call", w)) < REPLAY_TOL or \
        gpu.mc_eur(*P1, 1, "call", replay=w).price == oracle.mc_eur(*P1, 1, "call", w)
    w = oracle.normals_mt19937(4, math.sqrt(1 / 5), 5)
    assert abs(gpu.mc_amer(*P1, 2, 5, "put", replay=w).price - oracle.mc_amer(*P1, 2, 5, "put", w)) < 1e-10
    w = oracle.normals_mt19937(4, 1.0, 3)
    assert rel(gpu.mc_asia(*P1, 3, 1, "call", replay=w).price, oracle.mc_asia(*P1, 3, 1, "call", w)) < REPLAY_TOL or \
        gpu.mc_asia(*P1, 3, 1, "call", replay=w).price == 0.0
    # M = 1 American: no exercise dates before maturity
    w = oracle.normals_mt19937(4, 1.0, 500)
    assert rel(gpu.mc_amer(*P1, 1000, 1, "put", replay=w).price, oracle.mc_amer(*P1, 1000, 1, "put", w)) < REPLAY_TOL
    # tiny lattices
    for N in (1, 2, 3, 4, 15, 16, 17):
        for pf in ("call", "put"):
            g = gpu.binom(100, 95, .05, .3, 1, N, pf).price
            o = oracle.binom(100, 95, .05, .3, 1, N, pf)
            assert abs(g - o) < 1e-11 * max(1, abs(o)), (N, pf)


def test_replay_parity_on_random_parameters(gpu):
    # Beyond the golden vectors: seeded random parameter sets, every method through the C ABI in replay mode against the
    # oracle on the same mt19937 stream (1e-12 relative; the trees bit-equal). tests/test_oracle.py pins the oracle to
    # the compiled reference on random sets of the same kind.
    rng = np.random.default_rng(31415)
    for k in range(16):
        S0 = float(rng.uniform(60, 140)); E = float(rng.uniform(60, 140))
        r = float(rng.uniform(0.0, 0.1)); sigma = float(rng.uniform(0.05, 0.6)); T = float(rng.uniform(0.25, 2.0))
        pf = "call" if k % 2 else "put"
        seed = int(rng.integers(1, 10**6))
        N = int(rng.integers(100, 3000)) * 2
        M = int(rng.integers(2, 60))
        d = int(rng.integers(1, 17)); rho = float(rng.uniform(0.0, 0.9))
        tol = lambda want: REPLAY_TOL * max(abs(want), 1e-3)
        w = oracle.normals_mt19937(seed, math.sqrt(T), N)
        want = oracle.mc_eur(S0, E, r, sigma, T, N, pf, w)
        assert abs(gpu.mc_eur(S0, E, r, sigma, T, N, pf, replay=w).price - want) <= tol(want), (k, "eur")
        w = oracle.normals_mt19937(seed, math.sqrt(T / M), N * M)
        want = oracle.mc_asia(S0, E, r, sigma, T, N, M, pf, w)
        assert abs(gpu.mc_asia(S0, E, r, sigma, T, N, M, pf, replay=w).price - want) <= tol(want), (k, "asia")
        Z = oracle.normals_mt19937(seed, 1.0, N * d)
        want = oracle.mc_basket(S0, E, r, sigma, T, N, pf, d, rho, Z)
        assert abs(gpu.mc_eur_multi(S0, E, r, sigma, T, N, pf, d, rho, replay=Z).price - want) <= tol(want), (k, "basket", d)
        w = oracle.normals_mt19937(seed, math.sqrt(T / M), N // 2 * M)
        try:
            want = oracle.mc_amer(S0, E, r, sigma, T, N, M, pf, w)
        except Exception:   # singular regression: the reference throws (common.h:115-117), so does the library
            with pytest.raises(ValueError):
                gpu.mc_amer(S0, E, r, sigma, T, N, M, pf, replay=w)
        else:
            assert abs(gpu.mc_amer(S0, E, r, sigma, T, N, M, pf, replay=w).price - want) <= tol(want), (k, "amer")
        Nt = int(rng.integers(1, 3000))
        assert gpu.binom_vanilla_eur(S0, E, r, sigma, T, Nt, pf).price == oracle.binom_tree(S0, E, r, sigma, T, Nt, pf, False)
        assert gpu.binom_vanilla_amer(S0, E, r, sigma, T, Nt, pf).price == oracle.binom_tree(S0, E, r, sigma, T, Nt, pf, True)
        want = oracle.binom(S0, E, r, sigma, T, Nt, pf)
        assert abs(gpu.binom(S0, E, r, sigma, T, Nt, pf).price - want) <= 1e-10 * max(abs(want), 1e-3), (k, "binom")


def test_error_behaviour(gpu):
    with pytest.raises(ValueError, match="divisible by 2"):
        gpu.mc_amer(*P1, 1001, 10, "put")                      # reference include/common.h:180
    with pytest.raises(ValueError, match="Unknown payoff"):
        gpu.mc_asia(*P1, 100, 10, "digital")                   # reference src/mc_asia.cpp:56
    with pytest.raises(ValueError):
        gpu.mc_eur_multi(*P1, 100, "call", 33, 0.5)            # beyond PCF_MAX_ASSETS
    with pytest.raises(ValueError, match="positive definite"):
        gpu.mc_eur_multi(*P1, 100, "call", 4, -0.5)
    with pytest.raises(ValueError):
        gpu.mc_eur(*P1, 100, "call", replay=np.zeros(50))      # replay stream too short
    with pytest.raises(ValueError):
        gpu.mc_eur(*P1, 0, "call")
    # the limits INTEGRATION.md lists (the reference takes any int): refused, never truncated
    with pytest.raises(ValueError):
        gpu.mc_amer(*P1, 1000, 2049, "put")                    # exercise dates beyond the discount tables
    assert gpu.mc_amer(*P1, 64, 2048, "put", seed=1).price > 0    # the last M that is accepted (uint16 dates)
    with pytest.raises(ValueError):
        gpu.binom_vanilla_amer(*P1, 10_000_001, "put")         # O(N^2) tree beyond the documented cap
    with pytest.raises(ValueError):
        gpu.binom(*P1, 2 ** 31, "call")                        # the reference's int N
    # parameter sets that would drive exp() out of its finite range (the reference prints inf / NaN there) are refused,
    # not priced wrongly: sigma = 90 puts |x| up to 90*8.5 in play; so does a replayed "normal" of 1e6
    for fn, extra in ((gpu.mc_eur, ()), (gpu.mc_asia, (1,)), (gpu.mc_amer, (1,))):
        with pytest.raises(ValueError):
            fn(100, 100, .05, 90.0, 1, 1000, *extra, "call")
    with pytest.raises(ValueError):
        gpu.mc_eur_multi(100, 100, .05, 90.0, 1, 1000, "call", 4, 0.5)
    with pytest.raises(ValueError):
        gpu.mc_eur(*P1, 4, "call", replay=np.array([0.1, 1e6, 0.0, float("nan")]))
    assert gpu.mc_eur(100, 100, .05, 3.0, 1, 1000, "call", seed=1).price > 0      # large but representable: priced
    # more GPUs than the box has: an error at init (reference src/mc_eur_mpi.cpp:58-62 fails loudly at MPI_Init), not a clamp
    gpu.shutdown()
    try:
        with pytest.raises(ValueError):
            gpu.init(1000)
    finally:
        gpu.init(1)
    # a path store that cannot fit (2e9 paths x 50 dates = 800 GB) is refused cleanly and the library stays usable
    with pytest.raises(Exception) as ei:
        gpu.mc_amer(*P1, 2_000_000_000, 50, "put")
    assert getattr(ei.value, "status", None) == gpu.PCF_ENOMEM
    assert gpu.mc_amer(*P1, 10_000, 10, "put", seed=1).price > 0
    # all paths identical (sigma = 0) with S > E: x'x is singular -> the reference throws (common.h:115-117)
    with pytest.raises(ValueError, match="Detereminant"):
        gpu.mc_amer(100, 90, .05, 0.0, 1, 1000, 10, "call")


def test_multi_gpu_in_process_matches_single(gpu):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    one = gpu.mc_asia(*P1, 1_000_001, 252, "call", seed=31)
    am1 = gpu.mc_amer(*P1, 1_000_002, 50, "put", seed=31)
    bn1 = gpu.binom(*P1, 1_000_001, "call")
    gpu.shutdown()
    gpu.init(2)
    try:
        two = gpu.mc_asia(*P1, 1_000_001, 252, "call", seed=31)
        am2 = gpu.mc_amer(*P1, 1_000_002, 50, "put", seed=31)
        bn2 = gpu.binom(*P1, 1_000_001, "call")
        assert two.gpus == 2
        assert rel(two.price, one.price) < 1e-13 and rel(am2.price, am1.price) < 1e-13
        assert rel(bn2.price, bn1.price) < 1e-13
    finally:
        gpu.shutdown()
        gpu.init(1)
