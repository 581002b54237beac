"""GPU tests of the drop-in executables (bin/*): same argv, same CSV row as the reference programs."""
import math
import os
import subprocess

import pytest

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


def run(prog, *args, env=None):
    e = dict(os.environ, **(env or {}))
    p = subprocess.run([os.path.join(BIN, prog), *map(str, args)], env=e, capture_output=True, text=True)
    return p.returncode, p.stdout.strip(), p.stderr


def test_row_format_matches_reference_layout(tmp_path):
    # reference row: include/common.h:212-249 -- 16 comma-separated fields, setprecision(10)
    N, M = 20000, 252
    w = oracle.normals_mt19937(42, math.sqrt(1 / M), N * M)
    f = tmp_path / "w.f64"
    w.tofile(f)
    rc, out, err = run("mc_asia", "call", 100, 100, 0.05, 0.2, 1, N, M, 1, env={"PCF_REPLAY": str(f), "PCF_COMPARISON": "5.5"})
    assert rc == 0, err
    fields = out.split(",")
    assert len(fields) == 16 and len(out.splitlines()) == 1
    assert fields[:11] == ["CUDA", "call", "100", "100", "0.05", "0.2", "1", str(N), str(M), "1", "1"]
    want = oracle.mc_asia(100, 100, .05, .2, 1, N, M, "call", w)
    assert fields[13] == f"{want:.10g}"
    assert fields[14] == f"{abs(want - 5.5):.10g}" and fields[15] == f"{want - 5.5:.10g}"
    if oracle.have_ref():
        ref = oracle.ref_row("mc_asia", "call", 100, 100, 0.05, 0.2, 1, N, M, seed=42)
        assert ref[1:9] == fields[1:9] and ref[10] == fields[10] and ref[13] == fields[13]


@pytest.mark.parametrize("prog,args,nfield", [
    ("mc_eur", ("put", 100, 100, 0.05, 0.2, 1, 100000), ("0", "1")),
    ("mc_amer", ("put", 100, 100, 0.05, 0.2, 1, 100000, 50), ("50", "1")),
    ("mc_eur_multi", ("call", 100, 100, 0.05, 0.2, 1, 100000, 16, 0.5), ("0", "16")),
    ("binom_embar", ("call", 100, 110, 0.02, 0.75, 1, 1000), ("0", "1")),
    ("binom_vanilla_eur", ("call", 100, 110, 0.02, 0.75, 1, 1000), ("0", "1")),
    ("binom_vanilla_amer", ("put", 100, 100, 0.05, 0.2, 1, 2000), ("0", "1")),
])
def test_every_front_end_prints_one_row(prog, args, nfield):
    rc, out, err = run(prog, *args, env={"PCF_SEED": "7"})
    assert rc == 0, err
    fields = out.split(",")
    assert len(fields) == 16 and fields[0] == ("CUDA_vanilla" if "vanilla" in prog else "CUDA") and fields[1] == args[0]
    assert (fields[8], fields[10]) == nfield
    assert math.isfinite(float(fields[13]))
    if prog in ("binom_embar", "binom_vanilla_eur"):
        assert fields[13] == "26.60882645"  # reference results/results_binom_embar.csv, N = 1000 (both rows)
    if prog == "binom_vanilla_amer":
        assert fields[13] == "6.090232044"  # oracle/_ref/binom_vanilla_amer put 100 100 0.05 0.2 1 2000


def test_error_behaviour_matches_reference():
    rc, out, err = run("mc_eur", "straddle", 100, 100, 0.05, 0.2, 1, 1000)
    assert rc != 0 and out == "" and "Unknown payoff function" in err      # src/mc_eur.cpp:42
    rc, out, err = run("mc_amer", "put", 100, 100, 0.05, 0.2, 1, 1001, 10)
    assert rc != 0 and out == "" and "divisible by 2" in err              # include/common.h:180


def test_comparison_header_is_read_at_run_time(tmp_path):
    # the reference's run-scripts rewrite include/comparison.h and rebuild (runscript_mc_eur.sh:22-25); the CUDA front
    # ends read that file from the working directory instead
    (tmp_path / "include").mkdir()
    (tmp_path / "include" / "comparison.h").write_text("#pragma once\ndouble comparison = 26.61224;\n")
    p = subprocess.run([os.path.join(BIN, "binom_embar"), "call", "100", "110", "0.02", "0.75", "1", "1000"],
                       cwd=tmp_path, capture_output=True, text=True)
    f = p.stdout.strip().split(",")
    assert p.returncode == 0 and f[13] == "26.60882645"
    assert abs(float(f[15]) - (float(f[13]) - 26.61224)) < 1e-8   # Error = Result - comparison (common.h:246), %.10g
    assert float(f[14]) == abs(float(f[15]))


def test_runscript_cuda_appends_reference_format_rows(tmp_path):
    # the CUDA leg of runscript.sh -- same header, same sweep parameters, one CUDA row per (N, gpus)
    env = dict(os.environ, RESULTS_DIR=str(tmp_path), PCF_NS="10000 80000", PCF_SEED="3")
    p = subprocess.run(["bash", os.path.join(ROOT, "runscript_cuda.sh"), "all"], cwd=ROOT, env=env,
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    header = "Method,Payoff,S0,E,r,sigma,T,N,M,Parallel,Nr_of_assets,T_overall,T_calculation,Result,Abs_Error,Error"
    want = {"binom_embar": 4, "mc_eur": 2, "mc_eur_multi": 2, "mc_amer": 2, "mc_asia": 2}
    for method, nrows in want.items():
        lines = (tmp_path / f"results_{method}.csv").read_text().strip().splitlines()
        assert lines[0] == header                                          # runscript_mc_eur.sh:13
        rows = [l.split(",") for l in lines[1:]]
        assert len(rows) == nrows and all(len(r) == 16 for r in rows)
        assert all(r[0] in ("CUDA", "CUDA_vanilla") and r[1] == "call" and r[2] == "100" for r in rows)
        assert [r[7] for r in rows if r[0] == "CUDA"] == ["10000", "80000"]
    amer = [l.split(",") for l in (tmp_path / "results_mc_amer.csv").read_text().strip().splitlines()[1:]]
    # comparison was baked from the American tree at N = 10000: Error = Result - comparison, |.| in Abs_Error
    assert all(abs(float(r[14]) - abs(float(r[15]))) < 1e-12 and float(r[14]) < 2.0 for r in amer)
    assert all(r[8] == "200" for r in amer)
