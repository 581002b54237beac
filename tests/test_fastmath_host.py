"""CPU test: parcompfin_b200/csrc/fastmath.cuh is __host__ __device__; sweep it against long double libm."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fastmath_accuracy_on_host(tmp_path):
    exe = str(tmp_path / "fm_test")
    subprocess.run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O2", "-std=c++17", "-mfma", "-x", "c++",
                    os.path.join(ROOT, "tests", "fastmath_host_test.cpp"),
                    os.path.join(ROOT, "parcompfin_b200", "csrc", "fastmath_tables.cpp"), "-o", exe], check=True)
    out = subprocess.check_output([exe]).decode()
    v = {k: float(x) for k, x in (line.split() for line in out.strip().splitlines())}
    # -2 ln u: 1 ulp of a value <= 73 (absolute 1.4e-14 at the far tail), strictly positive at u = 1
    assert v["neg2log_abs_err"] < 2e-14 and v["neg2log_min"] > 0
    assert v["sqrt_ulp"] <= 0.51
    assert v["sincos_abs_err"] < 3e-16 and v["sincos_norm_err"] < 1e-15
    assert v["exp_small_ulp"] < 1.5 and v["exp_small_pm_ulp"] < 2.0 and v["exp_table_ulp"] < 1.5


def test_binomial_pair_math_on_host_matches_exact_sums(tmp_path, golden):
    """binom_math.cuh (host build, -ffp-contract=off like the library: the lattice parameters are
    cancellation-sensitive) summed on the CPU against the 50-digit exact sums and the reference."""
    exe = str(tmp_path / "binom_host")
    subprocess.run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O2", "-std=c++17", "-mfma",
                    "-ffp-contract=off", "-x", "c++", os.path.join(ROOT, "tests", "binom_host_test.cpp"),
                    os.path.join(ROOT, "parcompfin_b200", "csrc", "fastmath_tables.cpp"), "-o", exe], check=True)
    exact = [c for c in golden["exact_binom"]["cases"] if c["N"] <= 10 ** 6]
    ref = [c for c in golden["reference_vectors"]["binom_embar"] if c["N"] <= 32000]
    args = []
    for c in exact + ref:
        args += [c["payoff"]] + [repr(float(x)) for x in c["params"]] + [str(c["N"])]
    out = [float(x) for x in subprocess.check_output([exe] + args).decode().split()]
    for c, o in zip(exact, out[:len(exact)]):
        assert abs(o - c["price"]) <= 1e-14 * c["price"], c
    for c, o in zip(ref, out[len(exact):]):
        assert abs(o - c["price"]) <= 1e-10 * c["price"], c
