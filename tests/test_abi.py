"""CPU tests: the C-ABI library loads, exports every symbol include/pcf.h declares, mirrors the
reference's error behaviour in the host layer, and fails loudly without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest

import parcompfin_b200 as pcf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pcf.h")).read()
    return sorted(set(re.findall(r"PCF_API\s+[\w\s\*]+?\b(pcf_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = pcf.load_library()
    syms = header_symbols()
    assert len(syms) >= 18
    assert sorted(pcf.EXPORTS) == syms
    for s in syms:
        assert hasattr(lib, s), s


def test_struct_layout_matches_header():
    # sizes the C compiler produces for the two public structs (x86-64 SysV)
    assert ctypes.sizeof(pcf.PcfParams) == 104
    assert ctypes.sizeof(pcf.PcfResult) == 80


def test_strerror_mirrors_reference_messages():
    lib = pcf.load_library()
    assert lib.pcf_strerror(pcf.PCF_EINVAL_PAYOFF) == b"Unknown payoff function"          # src/mc_eur.cpp:42
    assert lib.pcf_strerror(pcf.PCF_EODD_N) == b"N needs to be divisible by 2 for finding paths"  # common.h:180
    assert lib.pcf_strerror(pcf.PCF_ESINGULAR) == b"Detereminant is not > 0"              # common.h:116 (sic)


def test_unknown_payoff_raises_like_reference():
    with pytest.raises(ValueError, match="Unknown payoff function"):
        pcf.mc_eur(100, 100, .05, .2, 1, 10, "straddle")


def test_shard_partition_covers_units_exactly():
    for units in (0, 1, 7, 8, 1000, 10**9 + 7):
        for world in (1, 2, 3, 4, 8):
            parts = [pcf.shard_of(units, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == units
            for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
                assert e0 == b1 and b0 <= e0
            assert sum(e - b for b, e in parts) == units


def test_chol_equicorr_is_host_side():
    import numpy as np
    L = pcf.chol_equicorr(16, 0.5)
    assert np.allclose(L @ L.T, 0.5 + 0.5 * np.eye(16), atol=1e-14)
    with pytest.raises(ValueError):
        pcf.chol_equicorr(3, -0.9)


def test_normal_transform_is_host_side():
    # SURVEY 8f.4 / include/mvn.h:63-76: Cholesky when positive definite, eigenvectors * sqrt(eigenvalues) otherwise
    import numpy as np
    rng = np.random.default_rng(1)
    B = rng.standard_normal((12, 12))
    cov = B @ B.T / 12 + 0.1 * np.eye(12)
    A, eig = pcf.normal_transform(cov)
    assert not eig and np.allclose(A, np.linalg.cholesky(cov), atol=1e-13) and np.allclose(np.triu(A, 1), 0)
    # rank-deficient (positive SEMI-definite): rho = 1 and rho = -1/(d-1)
    for d, rho in [(5, 1.0), (4, -1.0 / 3.0), (32, 1.0)]:
        cov = (1 - rho) * np.eye(d) + rho * np.ones((d, d))
        A, eig = pcf.normal_transform(cov)
        assert eig and np.allclose(A @ A.T, cov, atol=1e-12), (d, rho)
    B = rng.standard_normal((9, 3))
    cov = B @ B.T  # rank 3 of 9
    A, eig = pcf.normal_transform(cov)
    assert eig and np.allclose(A @ A.T, cov, atol=1e-12)
    with pytest.raises(ValueError):  # a negative eigenvalue is not a covariance matrix
        pcf.normal_transform(np.array([[1.0, 2.0], [2.0, 1.0]]))
    with pytest.raises(ValueError):  # asymmetric
        pcf.normal_transform(np.array([[1.0, 0.5], [0.2, 1.0]]))
    assert ctypes.sizeof(pcf.PcfBasket) == 40


def test_header_is_plain_c_and_links(tmp_path):
    """include/pcf.h compiled as C99 -pedantic by a consumer that links libpcf.so: struct sizes match the ctypes mirror,
    host-only entry points work, compute before pcf_init() is PCF_ENOINIT, a bad payoff is the reference's message."""
    import subprocess
    exe = str(tmp_path / "abi_c_test")
    libdir = os.path.join(ROOT, "parcompfin_b200")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_c_test.c"), "-o", exe, "-L", libdir, "-lpcf",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.check_output([exe]).decode().splitlines()
    assert out[0] == f"sizeof {ctypes.sizeof(pcf.PcfParams)} {ctypes.sizeof(pcf.PcfResult)} {ctypes.sizeof(pcf.PcfBasket)} abi 1"
    assert out[1].startswith("chol 0 ") and abs(float(out[1].split()[2]) - pcf.chol_equicorr(16, 0.5)[15, 15]) < 1e-16
    assert out[2].startswith("transform 0 0 ") and abs(float(out[2].split()[3]) - 0.75 ** 0.5) < 1e-15
    assert out[3] == "noinit 12 pcf_init() has not been called"
    assert out[4] == "badpayoff 1 Unknown payoff function"


def test_compute_without_init_or_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(pcf.PcfError):
        pcf.init(1)
    with pytest.raises(pcf.PcfError):
        pcf.mc_asia(100, 100, .05, .2, 1, 100, 10, "call")
