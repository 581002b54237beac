"""oracle -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so (oracle/cpu_ref.cpp),
the CPU restatement of the reference's hot path, plus helpers to run the compiled reference
binaries under oracle/_ref/. Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs only; the product package never imports it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_d, _ll, _i, _u64, _u32 = ctypes.c_double, ctypes.c_longlong, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32
_P = ctypes.POINTER(_d)
_lib = None


def build(ref: bool = True) -> None:
    """make -C oracle port [ref]; `ref` is skipped silently when /root/reference is absent."""
    subprocess.run(["make", "-C", HERE, "port"], check=True, stdout=subprocess.DEVNULL)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, stdout=subprocess.DEVNULL)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        L = ctypes.CDLL(LIB_PATH)
        L.oracle_normals_mt19937.argtypes = [_u64, _d, _ll, _P]
        L.oracle_mc_eur.restype = _d
        L.oracle_mc_eur.argtypes = [_d] * 5 + [_ll, _i, _P, _P, _P]
        L.oracle_mc_asia.restype = _d
        L.oracle_mc_asia.argtypes = [_d] * 5 + [_ll, _i, _i, _P, _P, _P]
        L.oracle_pathsfinder.restype = _i
        L.oracle_pathsfinder.argtypes = [_d] * 4 + [_ll, _i, _P, _P]
        L.oracle_mc_amer.restype = _d
        L.oracle_mc_amer.argtypes = [_d] * 5 + [_ll, _i, _i, _P, ctypes.POINTER(_i)]
        L.oracle_mc_amer_lsm.restype = _d
        L.oracle_mc_amer_lsm.argtypes = [_d] * 5 + [_ll, _i, _i, _P, ctypes.POINTER(_i)]
        L.oracle_chol_equicorr.restype = _i
        L.oracle_chol_equicorr.argtypes = [_i, _d, _P]
        L.oracle_mvn_transform.argtypes = [_i, _d, _P, ctypes.POINTER(_i)]
        L.oracle_mc_basket.restype = _d
        L.oracle_mc_basket.argtypes = [_d] * 5 + [_ll, _i, _i, _d, _P, _P, _P, ctypes.POINTER(_i)]
        L.oracle_mc_basket_general.restype = _d
        L.oracle_mc_basket_general.argtypes = [_P, _d, _d, _P, _d, _ll, _i, _i, _P, _P, _P, _P, _P]
        L.oracle_mc_basket_omp_timed.restype = _d
        L.oracle_mc_basket_omp_timed.argtypes = [_d] * 5 + [_ll, _i, _i, _d, _u64, _i, _P]
        L.oracle_mc_asia_omp_timed.restype = _d
        L.oracle_mc_asia_omp_timed.argtypes = [_d] * 5 + [_ll, _i, _i, _u64, _i, _P]
        L.oracle_mc_eur_omp_timed.restype = _d
        L.oracle_mc_eur_omp_timed.argtypes = [_d] * 5 + [_ll, _i, _u64, _i, _P]
        L.oracle_binom_params.argtypes = [_d, _d, _d, _i, _P, _P, _P, _P]
        L.oracle_binom.restype = _d
        L.oracle_binom.argtypes = [_d] * 5 + [_i, _i, _i]
        L.oracle_binom_tree.restype = _d
        L.oracle_binom_tree.argtypes = [_d] * 5 + [_i, _i, _i]
        L.oracle_philox4x32_10.argtypes = [ctypes.POINTER(_u32)] * 3
        L.oracle_normal_stream.argtypes = [_u64, _u32, _u64, _ll, _i, _d, _P]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(_P)


def _cp(payoff_fun) -> int:
    return 1 if payoff_fun in ("call", 1, 1.0) else -1


def normals_mt19937(seed: int, sd: float, n: int) -> np.ndarray:
    """std::mt19937(seed) + std::normal_distribution<>{0,sd}: the reference's own stream."""
    out = np.empty(n, dtype=np.float64)
    lib().oracle_normals_mt19937(seed, sd, n, _p(out))
    return out


def mc_eur(S0, E, r, sigma, T, N, payoff_fun, w, moments=False):
    w = np.ascontiguousarray(w, dtype=np.float64)
    s, s2 = _d(), _d()
    price = lib().oracle_mc_eur(S0, E, r, sigma, T, N, _cp(payoff_fun), _p(w), ctypes.byref(s), ctypes.byref(s2))
    return (price, s.value, s2.value) if moments else price


def mc_asia(S0, E, r, sigma, T, N, M, payoff_fun, dB, moments=False):
    dB = np.ascontiguousarray(dB, dtype=np.float64)
    s, s2 = _d(), _d()
    price = lib().oracle_mc_asia(S0, E, r, sigma, T, N, M, _cp(payoff_fun), _p(dB), ctypes.byref(s), ctypes.byref(s2))
    return (price, s.value, s2.value) if moments else price


def pathsfinder(S0, r, sigma, T, N, M, w) -> np.ndarray:
    w = np.ascontiguousarray(w, dtype=np.float64)
    paths = np.empty((M + 1, N), dtype=np.float64)
    if lib().oracle_pathsfinder(S0, r, sigma, T, N, M, _p(w), _p(paths)):
        raise ValueError("N needs to be divisible by 2 for finding paths")
    return paths


def mc_amer(S0, E, r, sigma, T, N, M, payoff_fun, w, lsm=False):
    w = np.ascontiguousarray(w, dtype=np.float64)
    st = _i()
    fn = lib().oracle_mc_amer_lsm if lsm else lib().oracle_mc_amer
    price = fn(S0, E, r, sigma, T, N, M, _cp(payoff_fun), _p(w), ctypes.byref(st))
    if st.value == 1:
        raise ValueError("N needs to be divisible by 2 for finding paths")
    if st.value == 2:
        raise ValueError("Detereminant is not > 0")
    return price


def chol_equicorr(d: int, rho: float) -> np.ndarray:
    L = np.zeros((d, d), dtype=np.float64)
    if lib().oracle_chol_equicorr(d, rho, _p(L)):
        raise ValueError("not positive definite")
    return L


def mvn_transform(d: int, rho: float):
    """include/mvn.h:53-76: (normTransform row-major d x d, used_eigen). Cholesky factor, or eigenvectors *
    sqrt(eigenvalues) when LLT reports a non-positive pivot (may hold NaN, exactly as in the reference)."""
    A = np.zeros((d, d), dtype=np.float64)
    eig = _i()
    lib().oracle_mvn_transform(d, rho, _p(A), ctypes.byref(eig))
    return A, bool(eig.value)


def mc_basket(S0, E, r, sigma, T, N, payoff_fun, d, rho, Z, moments=False):
    """src/mc_eur_multi.cpp:6-35 + include/mvn.h, both branches of mvn.h:68-76 (NaN where the reference gives NaN)."""
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    s, s2, st = _d(), _d(), _i()
    price = lib().oracle_mc_basket(S0, E, r, sigma, T, N, _cp(payoff_fun), d, rho, _p(Z), ctypes.byref(s),
                                   ctypes.byref(s2), ctypes.byref(st))
    return (price, s.value, s2.value) if moments else price


def mc_basket_general(S0, E, r, sigma, T, N, payoff_fun, A, weights, Z, moments=False):
    """per-asset spots/vols/weights and a full d x d normal transform A (Bt = A Z)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    d = A.shape[0]
    S0 = np.ascontiguousarray(np.broadcast_to(np.asarray(S0, dtype=np.float64), (d,)))
    sg = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, dtype=np.float64), (d,)))
    w = np.ascontiguousarray(weights, dtype=np.float64)
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    s, s2 = _d(), _d()
    price = lib().oracle_mc_basket_general(_p(S0), E, r, _p(sg), T, N, _cp(payoff_fun), d, _p(A), _p(w), _p(Z),
                                           ctypes.byref(s), ctypes.byref(s2))
    return (price, s.value, s2.value) if moments else price


def mc_basket_omp_timed(S0, E, r, sigma, T, N, payoff_fun, d, rho, seed, threads):
    sec = _d()
    price = lib().oracle_mc_basket_omp_timed(S0, E, r, sigma, T, N, _cp(payoff_fun), d, rho, seed, threads,
                                             ctypes.byref(sec))
    return price, sec.value


def mc_asia_omp_timed(S0, E, r, sigma, T, N, M, payoff_fun, seed, threads):
    sec = _d()
    price = lib().oracle_mc_asia_omp_timed(S0, E, r, sigma, T, N, M, _cp(payoff_fun), seed, threads, ctypes.byref(sec))
    return price, sec.value


def mc_eur_omp_timed(S0, E, r, sigma, T, N, payoff_fun, seed, threads):
    sec = _d()
    price = lib().oracle_mc_eur_omp_timed(S0, E, r, sigma, T, N, _cp(payoff_fun), seed, threads, ctypes.byref(sec))
    return price, sec.value


def binom_params(r, sigma, T, N):
    u, d, p, q = _d(), _d(), _d(), _d()
    lib().oracle_binom_params(r, sigma, T, N, ctypes.byref(u), ctypes.byref(d), ctypes.byref(p), ctypes.byref(q))
    return u.value, d.value, p.value, q.value


def binom(S0, E, r, sigma, T, N, payoff_fun, threads=1):
    return lib().oracle_binom(S0, E, r, sigma, T, N, _cp(payoff_fun), threads)


def binom_tree(S0, E, r, sigma, T, N, payoff_fun, american=False):
    """binom_vanilla_eur / binom_vanilla_amer restated (backward induction, O(N^2))."""
    return lib().oracle_binom_tree(S0, E, r, sigma, T, N, _cp(payoff_fun), 1 if american else 0)


def philox4x32_10(ctr, key):
    c = (_u32 * 4)(*ctr)
    k = (_u32 * 2)(*key)
    o = (_u32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return tuple(o)


def normal_stream(seed, stream, index0, count, T, scale=1.0) -> np.ndarray:
    out = np.empty((count, T), dtype=np.float64)
    lib().oracle_normal_stream(seed, stream, index0, count, T, scale, _p(out))
    return out


# ---- the compiled, unmodified reference (oracle/_ref, built by oracle/Makefile `ref`) ----------------
def have_ref() -> bool:
    return os.path.exists(os.path.join(REF_DIR, ".stamp"))


def ref_fn(prog: str, *args, seed: int = 42) -> float:
    """Runs oracle/_ref/<prog>_fn (function-level harness, %.17g) with time() pinned to `seed`."""
    env = dict(os.environ, PCF_FIXED_TIME=str(seed))
    out = subprocess.check_output([os.path.join(REF_DIR, prog + "_fn"), *map(str, args)], env=env)
    return float(out)


def ref_row(prog: str, *args, seed: int | None = None, timeout: float | None = None) -> list[str]:
    """Runs oracle/_ref/<prog> (the reference executable) and returns its CSV row split on ','."""
    env = dict(os.environ)
    if seed is not None:
        env["PCF_FIXED_TIME"] = str(seed)
    out = subprocess.check_output([os.path.join(REF_DIR, prog), *map(str, args)], env=env, timeout=timeout)
    return out.decode().strip().split(",")
