"""parcompfin_b200 -- host-side binding of libpcf.so (include/pcf.h).

The reference (moledoc/parcompfin) exposes each method as a free function called from its own
``main`` (e.g. ``mc_asia(S0,E,r,sigma,T,N,M,payoff_fun)``, reference src/mc_asia.cpp:5-14). The
functions below keep those names, argument order and meaning (``payoff_fun`` is +1/-1 or
"call"/"put") and raise ``ValueError`` where the reference throws ``std::invalid_argument``.
All arithmetic happens in the CUDA library; there is no Python or CPU fallback -- if libpcf.so is
missing or no GPU is visible the call fails.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCF_LIB: another build of the same library (the `make lib TUNING=1` flavour with the launch-shape knobs, tools/tune_*.py)
LIB_PATH = os.environ.get("PCF_LIB") or os.path.join(_HERE, "libpcf.so")

# status codes (include/pcf.h)
PCF_OK, PCF_EINVAL_PAYOFF, PCF_EODD_N, PCF_ESINGULAR, PCF_EINVAL, PCF_ENOTPD = 0, 1, 2, 3, 4, 5
PCF_ECUDA, PCF_ENCCL, PCF_ENOINIT, PCF_ENOMEM = 10, 11, 12, 13
PCF_FLAG_BINOM_WINDOW = 0x1
PCF_FLAG_AMER_LSM = 0x2
PCF_FLAG_BINOM_NOSCREEN = 0x4
PCF_FLAG_BASKET_GENERAL = 0x8
PCF_FLAG_TREE_WARP = 0x10
STREAM_EUR, STREAM_ASIA, STREAM_BASKET, STREAM_AMER = 0, 1, 2, 3
MAX_ASSETS = 32

# every symbol include/pcf.h declares (tests check the library exports all of them)
EXPORTS = (
    "pcf_init", "pcf_init_rank", "pcf_nccl_unique_id", "pcf_shutdown", "pcf_world_size",
    "pcf_ipc_export", "pcf_ipc_import", "pcf_peer_enable", "pcf_peer_active",
    "pcf_mc_eur", "pcf_mc_eur_multi", "pcf_mc_asia", "pcf_mc_amer", "pcf_binom_embar",
    "pcf_binom_vanilla_eur", "pcf_binom_vanilla_amer", "pcf_mc_basket", "pcf_normal_transform",
    "pcf_normal_stream", "pcf_philox4x32_10", "pcf_chol_equicorr", "pcf_fp64_peak", "pcf_hbm_peak",
    "pcf_device_info", "pcf_strerror", "pcf_last_error",
)


class PcfParams(ctypes.Structure):
    _fields_ = [
        ("S0", ctypes.c_double), ("E", ctypes.c_double), ("r", ctypes.c_double),
        ("sigma", ctypes.c_double), ("T", ctypes.c_double),
        ("cp", ctypes.c_int), ("N", ctypes.c_longlong), ("M", ctypes.c_int),
        ("assets", ctypes.c_int), ("rho", ctypes.c_double),
        ("seed", ctypes.c_ulonglong),
        ("replay", ctypes.POINTER(ctypes.c_double)), ("replay_len", ctypes.c_longlong),
        ("flags", ctypes.c_uint),
    ]


class PcfBasket(ctypes.Structure):
    _fields_ = [(n, ctypes.POINTER(ctypes.c_double)) for n in ("S0", "sigma", "weight", "cov", "transform")]


class PcfResult(ctypes.Structure):
    _fields_ = [
        ("price", ctypes.c_double), ("sum", ctypes.c_double), ("sumsq", ctypes.c_double),
        ("std_error", ctypes.c_double), ("n", ctypes.c_longlong), ("units", ctypes.c_longlong),
        ("seconds_kernel", ctypes.c_double), ("seconds_total", ctypes.c_double),
        ("launches", ctypes.c_int), ("gpus", ctypes.c_int), ("status", ctypes.c_int),
    ]


@dataclass
class Result:
    price: float
    sum: float
    sumsq: float
    std_error: float
    n: int
    units: int
    seconds_kernel: float
    seconds_total: float
    launches: int
    gpus: int


class PcfError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"libpcf status {status}: {text}")
        self.status = status


_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen libpcf.so (in-tree, next to this file). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make lib` or `python -c 'import __graft_entry__ as g; "
            "g.build()'` (there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    P, R = ctypes.POINTER(PcfParams), ctypes.POINTER(PcfResult)
    for name in ("pcf_mc_eur", "pcf_mc_eur_multi", "pcf_mc_asia", "pcf_mc_amer", "pcf_binom_embar",
                 "pcf_binom_vanilla_eur", "pcf_binom_vanilla_amer"):
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = [P, R], ctypes.c_int
    lib.pcf_mc_basket.argtypes, lib.pcf_mc_basket.restype = [P, ctypes.POINTER(PcfBasket), R], ctypes.c_int
    lib.pcf_normal_transform.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
    lib.pcf_normal_transform.restype = ctypes.c_int
    lib.pcf_init.argtypes, lib.pcf_init.restype = [ctypes.c_int], ctypes.c_int
    lib.pcf_init_rank.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
    lib.pcf_init_rank.restype = ctypes.c_int
    lib.pcf_nccl_unique_id.argtypes, lib.pcf_nccl_unique_id.restype = [ctypes.c_char_p], ctypes.c_int
    lib.pcf_ipc_export.argtypes, lib.pcf_ipc_export.restype = [ctypes.c_char_p], ctypes.c_int
    lib.pcf_ipc_import.argtypes, lib.pcf_ipc_import.restype = [ctypes.c_char_p, ctypes.c_int], ctypes.c_int
    lib.pcf_peer_enable.argtypes, lib.pcf_peer_enable.restype = [ctypes.c_int], ctypes.c_int
    lib.pcf_peer_active.argtypes, lib.pcf_peer_active.restype = [], ctypes.c_int
    lib.pcf_shutdown.argtypes, lib.pcf_shutdown.restype = [], ctypes.c_int
    lib.pcf_world_size.argtypes, lib.pcf_world_size.restype = [], ctypes.c_int
    lib.pcf_normal_stream.argtypes = [ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_ulonglong,
                                      ctypes.c_longlong, ctypes.c_int, ctypes.c_double,
                                      ctypes.POINTER(ctypes.c_double)]
    lib.pcf_normal_stream.restype = ctypes.c_int
    lib.pcf_philox4x32_10.argtypes = [ctypes.POINTER(ctypes.c_uint)] * 3
    lib.pcf_philox4x32_10.restype = ctypes.c_int
    lib.pcf_chol_equicorr.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
    lib.pcf_chol_equicorr.restype = ctypes.c_int
    lib.pcf_fp64_peak.argtypes = [ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
    lib.pcf_fp64_peak.restype = ctypes.c_int
    lib.pcf_hbm_peak.argtypes = [ctypes.c_longlong, ctypes.POINTER(ctypes.c_double)]
    lib.pcf_hbm_peak.restype = ctypes.c_int
    lib.pcf_device_info.argtypes = [ctypes.c_char_p, ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 3 + \
        [ctypes.POINTER(ctypes.c_longlong)]
    lib.pcf_device_info.restype = ctypes.c_int
    lib.pcf_strerror.argtypes, lib.pcf_strerror.restype = [ctypes.c_int], ctypes.c_char_p
    lib.pcf_last_error.argtypes, lib.pcf_last_error.restype = [], ctypes.c_char_p
    _lib = lib
    return lib


def _check(status: int) -> None:
    if status == PCF_OK:
        return
    lib = load_library()
    text = lib.pcf_strerror(status).decode()
    if status in (PCF_EINVAL_PAYOFF, PCF_EODD_N, PCF_ESINGULAR, PCF_EINVAL, PCF_ENOTPD):
        # the reference throws std::invalid_argument at these points
        raise ValueError(text)
    detail = lib.pcf_last_error().decode()
    raise PcfError(status, f"{text} ({detail})" if detail else text)


# ---- lifetime -------------------------------------------------------------------------------------
def init(gpus: int = 1) -> None:
    """Single process driving the first ``gpus`` visible devices (0 = all)."""
    _check(load_library().pcf_init(gpus))


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    _check(load_library().pcf_nccl_unique_id(buf))
    return buf.raw


def init_rank(rank: int, world: int, device: int, nccl_id: bytes | None = None) -> None:
    """One process per GPU; ``nccl_id`` is rank 0's :func:`nccl_unique_id`, distributed by the caller."""
    _check(load_library().pcf_init_rank(rank, world, device, nccl_id))


def ipc_export() -> bytes:
    """CUDA IPC handle (64 bytes) of this rank's exchange mailbox."""
    buf = ctypes.create_string_buffer(64)
    _check(load_library().pcf_ipc_export(buf))
    return buf.raw


def ipc_import(handles: list[bytes]) -> None:
    """Maps every rank's mailbox (handles in rank order) for the NVLink peer-memory exchange."""
    blob = b"".join(handles)
    _check(load_library().pcf_ipc_import(blob, len(handles)))


def peer_enable(on: bool) -> None:
    _check(load_library().pcf_peer_enable(1 if on else 0))


def peer_active() -> bool:
    return bool(load_library().pcf_peer_active())


def shutdown() -> None:
    if _lib is not None:
        _lib.pcf_shutdown()


def world_size() -> int:
    return load_library().pcf_world_size()


# ---- the hot path -----------------------------------------------------------------------------------
def _cp(payoff_fun) -> int:
    if payoff_fun in ("call", 1, 1.0):
        return 1
    if payoff_fun in ("put", -1, -1.0):
        return -1
    raise ValueError("Unknown payoff function")  # reference src/mc_eur.cpp:42


def shard_of(units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition used by the library (csrc/common.cuh shard_of)."""
    per = (units + world - 1) // world
    b, e = min(per * rank, units), min(per * (rank + 1), units)
    return b, e


def _call(fn_name, S0, E, r, sigma, T, N, payoff_fun, M=0, assets=1, rho=0.0, seed=0, replay=None,
          flags=0, basket=None) -> Result:
    lib = load_library()
    p = PcfParams()
    p.S0, p.E, p.r, p.sigma, p.T = float(S0), float(E), float(r), float(sigma), float(T)
    p.cp, p.N, p.M, p.assets, p.rho = _cp(payoff_fun), int(N), int(M), int(assets), float(rho)
    p.seed, p.flags = int(seed) & 0xFFFFFFFFFFFFFFFF, int(flags)
    keep = None
    if replay is not None:
        keep = np.ascontiguousarray(replay, dtype=np.float64)
        p.replay = keep.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        p.replay_len = keep.size
    out = PcfResult()
    if basket is not None:
        _check(getattr(lib, fn_name)(ctypes.byref(p), ctypes.byref(basket), ctypes.byref(out)))
    else:
        _check(getattr(lib, fn_name)(ctypes.byref(p), ctypes.byref(out)))
    return Result(out.price, out.sum, out.sumsq, out.std_error, out.n, out.units, out.seconds_kernel,
                  out.seconds_total, out.launches, out.gpus)


def mc_eur(S0, E, r, sigma, T, N, payoff_fun, *, seed=0, replay=None) -> Result:
    """reference src/mc_eur.cpp:5-27"""
    return _call("pcf_mc_eur", S0, E, r, sigma, T, N, payoff_fun, seed=seed, replay=replay)


def mc_eur_multi(S0, E, r, sigma, T, N, payoff_fun, assets, rho, *, seed=0, replay=None, general=False) -> Result:
    """reference src/mc_eur_multi.cpp:6-35 (the function is also called mc_eur there). ``general=True``
    (PCF_FLAG_BASKET_GENERAL) prices through the general triangular kernel instead of the equicorrelation fast path."""
    return _call("pcf_mc_eur_multi", S0, E, r, sigma, T, N, payoff_fun, assets=assets, rho=rho,
                 seed=seed, replay=replay, flags=PCF_FLAG_BASKET_GENERAL if general else 0)


def mc_basket(S0, E, r, sigma, T, N, payoff_fun, assets, *, rho=0.0, weights=None, cov=None, transform=None,
              seed=0, replay=None) -> Result:
    """General basket (include/pcf.h pcf_mc_basket): ``S0`` and ``sigma`` may be scalars (the reference's
    case, src/mc_eur_multi.cpp:30) or length-``assets`` arrays; ``weights`` defaults to 1/d; ``cov`` is the d x d
    covariance of the driving normals (default: equicorrelation ``rho``, include/mvn.h:55-60) and ``transform`` an
    explicit factor A (Bt = A Z) that overrides it."""
    d = int(assets)
    keep = []

    def arr(x, shape):
        if x is None:
            return None
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), shape))
        if a.shape != shape:
            raise ValueError("basket array has the wrong shape")
        keep.append(a)
        return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

    b = PcfBasket()
    b.S0 = arr(S0, (d,)) if np.ndim(S0) else None
    b.sigma = arr(sigma, (d,)) if np.ndim(sigma) else None
    b.weight, b.cov, b.transform = arr(weights, (d,)), arr(cov, (d, d)), arr(transform, (d, d))
    s0 = float(np.ravel(S0)[0])
    sg = float(np.ravel(sigma)[0])
    return _call("pcf_mc_basket", s0, E, r, sg, T, N, payoff_fun, assets=d, rho=rho, seed=seed, replay=replay,
                 basket=b)


def normal_transform(cov) -> tuple[np.ndarray, bool]:
    """The factor A (A A^T = cov) the basket kernel uses, and whether the eigen fallback of mvn.h:72-76 produced it."""
    cov = np.ascontiguousarray(cov, dtype=np.float64)
    d = cov.shape[0]
    A = np.zeros((d, d), dtype=np.float64)
    eig = ctypes.c_int()
    P = ctypes.POINTER(ctypes.c_double)
    _check(load_library().pcf_normal_transform(d, cov.ctypes.data_as(P), A.ctypes.data_as(P), ctypes.byref(eig)))
    return A, bool(eig.value)


def mc_asia(S0, E, r, sigma, T, N, M, payoff_fun, *, seed=0, replay=None) -> Result:
    """reference src/mc_asia.cpp:5-40"""
    return _call("pcf_mc_asia", S0, E, r, sigma, T, N, payoff_fun, M=M, seed=seed, replay=replay)


def mc_amer(S0, E, r, sigma, T, N, M, payoff_fun, *, seed=0, replay=None, lsm=False) -> Result:
    """reference src/mc_amer.cpp:5-114. ``lsm=True`` switches the exercise rule to textbook Longstaff-Schwartz
    (PCF_FLAG_AMER_LSM); the default reproduces the reference's own rule."""
    return _call("pcf_mc_amer", S0, E, r, sigma, T, N, payoff_fun, M=M, seed=seed, replay=replay,
                 flags=PCF_FLAG_AMER_LSM if lsm else 0)


def binom(S0, E, r, sigma, T, N, payoff_fun, *, window=False, screen=True) -> Result:
    """reference src/binom_embar.cpp:5-50. ``screen=False`` (PCF_FLAG_BINOM_NOSCREEN) sends every term pair through
    the full-accuracy routine; the result is bit-identical, only slower."""
    return _call("pcf_binom_embar", S0, E, r, sigma, T, N, payoff_fun,
                 flags=(PCF_FLAG_BINOM_WINDOW if window else 0) | (0 if screen else PCF_FLAG_BINOM_NOSCREEN))


def binom_vanilla_eur(S0, E, r, sigma, T, N, payoff_fun, *, warp_tiling=False) -> Result:
    """reference src/binom_vanilla_eur.cpp:5-41 (backward-induction tree; `units` = node updates).
    ``warp_tiling=True`` (PCF_FLAG_TREE_WARP): the warp-trapezoid kernel instead of the CTA-cooperative one."""
    return _call("pcf_binom_vanilla_eur", S0, E, r, sigma, T, N, payoff_fun,
                 flags=PCF_FLAG_TREE_WARP if warp_tiling else 0)


def binom_vanilla_amer(S0, E, r, sigma, T, N, payoff_fun, *, warp_tiling=False) -> Result:
    """reference src/binom_vanilla_amer.cpp:5-42 (American tree; `units` = node updates)"""
    return _call("pcf_binom_vanilla_amer", S0, E, r, sigma, T, N, payoff_fun,
                 flags=PCF_FLAG_TREE_WARP if warp_tiling else 0)


# ---- diagnostics ---------------------------------------------------------------------------------
def normal_stream(seed: int, stream: int, index0: int, count: int, T: int, scale: float = 1.0) -> np.ndarray:
    out = np.empty((count, T), dtype=np.float64)
    _check(load_library().pcf_normal_stream(seed, stream, index0, count, T, scale,
                                            out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    return out


def philox4x32_10(ctr, key) -> tuple[int, int, int, int]:
    c = (ctypes.c_uint * 4)(*ctr)
    k = (ctypes.c_uint * 2)(*key)
    o = (ctypes.c_uint * 4)()
    _check(load_library().pcf_philox4x32_10(c, k, o))
    return tuple(o)


def chol_equicorr(d: int, rho: float) -> np.ndarray:
    L = np.zeros((d, d), dtype=np.float64)
    _check(load_library().pcf_chol_equicorr(d, rho, L.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    return L


def fp64_peak(seconds: float = 0.2) -> float:
    """Measured thread-level DFMA/s of the context's GPU (x2 = FP64 flop/s)."""
    v = ctypes.c_double()
    _check(load_library().pcf_fp64_peak(seconds, ctypes.byref(v)))
    return v.value


def hbm_peak(nbytes: int = 1 << 30) -> float:
    v = ctypes.c_double()
    _check(load_library().pcf_hbm_peak(nbytes, ctypes.byref(v)))
    return v.value


def device_info() -> dict:
    name = ctypes.create_string_buffer(256)
    sm, maj, mnr, mem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong()
    _check(load_library().pcf_device_info(name, 256, ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr),
                                          ctypes.byref(mem)))
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mnr.value),
            "mem_bytes": mem.value}
