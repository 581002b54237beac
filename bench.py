#!/usr/bin/env python
"""bench.py -- simulated path-steps per second of the Monte Carlo hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU)

A "step" is one pass of the hot path over one batch of synthetic parameters: one C-ABI call pricing the
whole workload (default: BASELINE.json config 3, mc_asia call 100/100/.05/.2/1, 1e9 paths x 252 dates,
split over the N GPUs by path index -- strong scaling, Philox keyed by global path index).

  value   path-steps/s from CUDA-event time on the launching stream (kernel + all-reduce), max over ranks
  e2e     same metric from the host wall clock around the C-ABI calls (parameter upload, launch, result
          read-back all inside), barrier + device sync on both sides, max over ranks
  roofline  algorithmic FP64-pipe slots (55 per path-step) / kernel time, against the DFMA
          issue peak measured in this run by a register-resident FMA-chain kernel (MEASURED_PEAKS.json
          has no FP64 figure)
  cpu_baseline  the reference's own OpenMP program (oracle/_ref/mc_asia_omp, compiled from the unmodified
          sources) on this box's host cores, bounded sample
  others  the other four BASELINE.json configs, one short measurement each, same definitions; configs 1 and 2 also at
          their STATED sizes (1e7 paths; N = 1e5, 1e6, 1e7), which are launch-latency bound: microseconds per call on the
          device clock and on the C-ABI call's own wall clock, next to the floor (a one-unit call of the same method)
  parity  (N > 1 only, outside the timed region) every method priced at a small size by the N ranks together -- with the
          NVLink peer mailboxes and again with ncclAllReduce -- against rank 0 alone on one GPU: relative differences and
          an overall `ok` (bound 1e-13; identical normal streams, only the summation order differs). The reference's own
          distributed check is `mpirun -n 4` in its *_tst targets (reference Makefile:45-46,79-80).

`--impl reference` times the reference's CPU implementation (rank 0 only) on a bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...", printed to stdout whenever NCCL_DEBUG is
# set on the box) goes to stderr instead. Set before anything loads libnccl.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

P = dict(S0=100.0, E=100.0, r=0.05, sigma=0.2, T=1.0)
SEED0 = 20240229

# algorithmic work per unit. FP64 "slots" = DFMA/DMUL/DADD issue slots; FMA-equivalent
# flops = 2 x slots.
WORKLOADS = {
    "mc_asia": dict(config="mc_asia call 100/100/.05/.2/1, 1e9 paths x 252 dates (BASELINE config 3)",
                    N=1_000_000_000, M=252, steps_per_unit=252, slots=55.0, bound="fp64", unit="path-steps/s",
                    kernel="mc_asia_kernel", exec_cycles=97.0,
                    exec_src="SASS of the shipped loop (tools/sass_count.py): 31.5 FP64 + 8.5 IMAD.WIDE per path-step"),
    "mc_eur": dict(config="mc_eur call 100/100/.05/.2/1, 2e9 paths (config 1 at the size SURVEY 8d quotes the "
                          "roofline on; 1e7 paths is 30 us of work)",
                   N=2_000_000_000, M=0, steps_per_unit=1, slots=60.0, bound="fp64", unit="path-steps/s",
                   kernel="mc_eur_kernel", exec_cycles=106.75,
                   exec_src="SASS of the shipped loop (tools/sass_count.py): 34.4 FP64 + 9.5 IMAD.WIDE per path"),
    "mc_eur_multi": dict(config="mc_eur_multi call, d=16 rho=0.5, 1e9 paths (BASELINE config 4)",
                         N=1_000_000_000, M=0, steps_per_unit=1, slots=990.0, bound="fp64", unit="path-steps/s",
                         kernel="mc_basket_equi_kernel", assets=16, rho=0.5),
    # SURVEY 8(f).4 (widening row): per-asset spots/vols/weights and a dense SPD covariance (no equicorrelation shortcut)
    "mc_basket_general": dict(config="mc_basket call, d=16, per-asset S0/sigma/weights, dense SPD covariance, 1e9 paths "
                                     "(SURVEY 8f.4: the general kernel, full triangular factor)",
                              N=1_000_000_000, M=0, steps_per_unit=1, slots=990.0, bound="fp64", unit="path-steps/s",
                              kernel="mc_basket_kernel", assets=16, exec_cycles=1920.0,
                              exec_src="ncu: 640 FP64 + 160 IMAD.WIDE warp instructions per path"),
    "mc_amer": dict(config="mc_amer put, 1e8 paths x 50 exercise dates, paths in HBM (BASELINE config 5)",
                    N=100_000_000, M=50, steps_per_unit=50, bytes=36.0, bound="hbm", unit="path-steps/s",
                    kernel="amer_sweep_kernel+amer_paths_kernel+amer_pad_kernel+amer_fill_when_kernel", traffic_all=True),
    "binom_embar": dict(config="binom_embar call 100/100/.05/.2/1, N=1e8 steps (BASELINE config 2)",
                        N=100_000_000, M=0, steps_per_unit=1, slots=85.0, bound="fp64", unit="terms/s",
                        kernel="binom_terms_kernel"),
    # the same sum with the screening pass off (PCF_FLAG_BINOM_NOSCREEN): every pair through the full-accuracy routine;
    # bit-identical result (tests/test_gpu_parity.py), reported so that the shortcut's share is visible
    "binom_embar_noscreen": dict(config="binom_embar call 100/100/.05/.2/1, N=1e8 steps, screening pass OFF "
                                        "(every term pair through the full-accuracy saddle-point routine)",
                                 N=100_000_000, M=0, steps_per_unit=1, slots=85.0, bound="fp64", unit="terms/s",
                                 kernel="binom_terms_kernel"),
    "binom_embar_max": dict(config="binom_embar call 100/100/.05/.2/1, N=2147483647 steps (the reference's int limit; SURVEY 8d "
                                   "quotes the binomial roofline at this size)",
                            N=2_147_483_647, M=0, steps_per_unit=1, slots=85.0, bound="fp64", unit="terms/s",
                            kernel="binom_terms_kernel"),
    # BASELINE configs 1 and 2 AS STATED (reference src/mc_eur.cpp:23-25 at N = 1e7, src/binom_embar.cpp:34-46 at N = 1e5..1e7):
    # tens of microseconds of work, bound by launch + completion latency (SURVEY 7 "Tiny workloads"); reported as
    # microseconds per call with the floor beside them, the roofline figures stay on the large variants above
    "mc_eur_stated": dict(config="mc_eur call 100/100/.05/.2/1, 1e7 paths (BASELINE config 1 as stated)",
                          N=10_000_000, M=0, steps_per_unit=1, slots=60.0, bound="fp64", unit="path-steps/s",
                          kernel="mc_eur_kernel", latency=True, method="mc_eur"),
    "binom_embar_1e5": dict(config="binom_embar call 100/100/.05/.2/1, N=1e5 steps (BASELINE config 2, smallest stated size)",
                            N=100_000, M=0, steps_per_unit=1, slots=85.0, bound="fp64", unit="terms/s",
                            kernel="binom_terms_kernel", latency=True, method="binom_embar"),
    "binom_embar_1e6": dict(config="binom_embar call 100/100/.05/.2/1, N=1e6 steps (BASELINE config 2)",
                            N=1_000_000, M=0, steps_per_unit=1, slots=85.0, bound="fp64", unit="terms/s",
                            kernel="binom_terms_kernel", latency=True, method="binom_embar"),
    "binom_embar_1e7": dict(config="binom_embar call 100/100/.05/.2/1, N=1e7 steps (BASELINE config 2)",
                            N=10_000_000, M=0, steps_per_unit=1, slots=85.0, bound="fp64", unit="terms/s",
                            kernel="binom_terms_kernel", latency=True, method="binom_embar"),
    # SURVEY 8(f).1 (widening row): algorithmic work = the recurrence as the reference writes it with tabulated powers:
    # 2 mul + add + IEEE division (10 slots, libdevice) [+ 2 mul, payoff 2, max 1 for the American tree]
    "binom_vanilla_amer": dict(config="binom_vanilla_amer put 100/100/.05/.2/1, N=1e5 layers (SURVEY 8f.1; the reference "
                                      "needs ~45 s per tree at this size)",
                               N=100_000, M=0, steps_per_unit=1, slots=18.0, bound="fp64", unit="node-updates/s",
                               algo_src="DESIGN 4.5: 2 mul + add + IEEE division 10 + 2 mul + payoff 2 + max 1",
                               kernel="tree_cta_kernel", replicas=True),
}


def run_ours_once(pcf, name, seed, N=None):
    w = WORKLOADS[name]
    N = N or w["N"]
    a = (P["S0"], P["E"], P["r"], P["sigma"], P["T"])
    if name == "mc_asia":
        return pcf.mc_asia(*a, N, w["M"], "call", seed=seed)
    if name in ("mc_eur", "mc_eur_stated"):
        return pcf.mc_eur(*a, N, "call", seed=seed)
    if name == "mc_eur_multi":
        return pcf.mc_eur_multi(*a, N, "call", w["assets"], w["rho"], seed=seed)
    if name == "mc_basket_general":
        import numpy as np
        rng = np.random.default_rng(16)
        d = w["assets"]
        B = rng.standard_normal((d, d))
        cov = B @ B.T / d + 0.2 * np.eye(d)
        return pcf.mc_basket(rng.uniform(80, 120, d), P["E"], P["r"], rng.uniform(.1, .4, d), P["T"], N, "call", d,
                             weights=rng.dirichlet(np.ones(d)), cov=cov, seed=seed)
    if name == "mc_amer":
        return pcf.mc_amer(*a, N, w["M"], "put", seed=seed)
    if name in ("binom_embar", "binom_embar_max", "binom_embar_1e5", "binom_embar_1e6", "binom_embar_1e7"):
        return pcf.binom(*a, N, "call")
    if name == "binom_embar_noscreen":
        return pcf.binom(*a, N, "call", screen=False)
    if name == "binom_vanilla_amer":
        return pcf.binom_vanilla_amer(*a, N, "put")
    raise KeyError(name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for ts, line in self.lines:
            if not (t0 <= ts <= t1 + 0.2):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_asia(n_paths, threads, seed=SEED0):
    """Times the reference's CPU implementation of the headline workload on `threads` host threads.
    Returns (path_steps_per_s, kind, seconds). kind = "reference" (oracle/_ref/mc_asia_omp, the unmodified
    source at -O2, its own T_calculation clock) or "port" (oracle/cpu_ref.cpp twin) when _ref is absent."""
    import oracle
    M = WORKLOADS["mc_asia"]["M"]
    if oracle.have_ref():
        row = oracle.ref_row("mc_asia_omp", "call", 100, 100, 0.05, 0.2, 1, n_paths, M, threads)
        sec = float(row[12])
        return n_paths * M / sec, "reference", sec
    oracle.mc_asia_omp_timed(100, 100, .05, .2, 1, 20000, M, "call", seed, threads)  # wake the thread pool
    _, sec = oracle.mc_asia_omp_timed(100, 100, .05, .2, 1, n_paths, M, "call", seed, threads)
    return n_paths * M / sec, "port", sec


def cpu_reference_other(name, threads):
    """cpu_baseline of the `others` entries (SURVEY 8d sizes, reduced so that each run takes a few seconds): the
    reference's own program from oracle/_ref on `threads` host threads, its T_calculation clock; the basket, which the
    reference cannot build here (Eigen/Boost absent), is the oracle restatement with the reference's OpenMP placement."""
    import oracle
    a = (100, 100, 0.05, 0.2, 1)
    try:
        if name == "mc_eur_multi":
            n = 2_000_000
            if oracle.have_ref() and os.path.exists(os.path.join(oracle.REF_DIR, "mc_eur_multi_omp")):
                sec = float(oracle.ref_row("mc_eur_multi_omp", "call", *a, n, 16, 0.5, threads)[12])
                return {"value": n / sec, "unit": "path-steps/s", "cores": threads, "kind": "reference",
                        "sample": f"{n} of 1e9 paths, d=16, {sec:.1f} s, mc_eur_multi_omp (unmodified reference source at "
                                  "-O2, compiled against the stand-in Eigen/Boost headers of oracle/shim; its sample "
                                  "generation is serial, src/mc_eur_multi_omp.cpp:31)"}
            _, sec = oracle.mc_basket_omp_timed(*a, n, "call", 16, 0.5, SEED0, threads)
            return {"value": n / sec, "unit": "path-steps/s", "cores": threads, "kind": "port",
                    "sample": f"{n} of 1e9 paths, d=16, {sec:.1f} s, oracle restatement with the OpenMP placement of "
                              "src/mc_eur_multi_omp.cpp:31-46 (serial sample generation)"}
        if not oracle.have_ref():
            return None
        if name == "mc_eur_stated":
            n = 10_000_000
            sec = float(oracle.ref_row("mc_eur_omp", "call", *a, n, threads)[12])
            return {"value": n / sec, "unit": "path-steps/s", "cores": threads, "kind": "reference",
                    "sample": f"the full 1e7 paths, {sec * 1e3:.1f} ms, mc_eur_omp (unmodified reference source, -O2)"}
        if name == "mc_eur":
            n = 400_000_000
            sec = float(oracle.ref_row("mc_eur_omp", "call", *a, n, threads)[12])
            return {"value": n / sec, "unit": "path-steps/s", "cores": threads, "kind": "reference",
                    "sample": f"{n} of 2e9 paths, {sec:.1f} s, mc_eur_omp (unmodified reference source, -O2)"}
        if name == "mc_amer":
            n, M = 2_000_000, 50
            sec = float(oracle.ref_row("mc_amer_omp", "put", *a, n, M, threads)[12])
            return {"value": n * M / sec, "unit": "path-steps/s", "cores": threads, "kind": "reference",
                    "sample": f"{n} of 1e8 paths x {M} dates, {sec:.1f} s, mc_amer_omp (unmodified reference source, -O2)"}
        if name in ("binom_embar", "binom_embar_noscreen", "binom_embar_max", "binom_embar_1e5", "binom_embar_1e6",
                    "binom_embar_1e7"):
            n = 50_000
            sec = float(oracle.ref_row("binom_embar_omp", "call", *a, n, threads)[12])
            return {"value": (n + 1) / sec, "unit": "terms/s", "cores": threads, "kind": "reference",
                    "sample": f"N={n} (the reference's comb() makes the sum O(N^2): terms/s falls as 1/N; N=1e8 is out of "
                              f"its reach), {sec:.1f} s, binom_embar_omp (unmodified reference source, -O2)"}
        if name == "binom_vanilla_amer":
            n = 20_000
            sec = float(oracle.ref_row("binom_vanilla_amer", "put", *a, n)[12])
            return {"value": n * (n + 1) / 2 / sec, "unit": "node-updates/s", "cores": 1, "kind": "reference",
                    "sample": f"N={n} of 1e5 layers, {sec:.1f} s, binom_vanilla_amer (unmodified reference source, -O2; the "
                              "reference has no parallel flavour of the tree)"}
    except Exception as ex:  # a missing binary must not take the GPU numbers down with it
        return {"error": str(ex)}
    return None


def bench_reference(args, rank):
    """--impl reference: rank 0 alone times the reference's CPU path; other ranks exit 0 without work."""
    if rank != 0:
        return
    threads = host_threads()
    w = WORKLOADS["mc_asia"]
    n_paths = args.ref_paths
    for _ in range(args.warmup):
        cpu_reference_asia(max(n_paths // 8, 1000), threads)
    secs, kind = [], "reference"
    for _ in range(args.steps):
        _, kind, sec = cpu_reference_asia(n_paths, threads)
        secs.append(sec)
    value = n_paths * w["M"] * args.steps / sum(secs)
    sample = (f"{n_paths} of 1e9 paths x {w['M']} dates per step (oracle/_ref/mc_asia_omp -O2, unmodified reference "
              f"source, {threads} OpenMP threads)" if kind == "reference" else
              f"{n_paths} of 1e9 paths x {w['M']} dates per step (oracle/cpu_ref.cpp OpenMP twin, {threads} threads)")
    emit({
        "impl": "reference", "metric": "path_steps_per_sec", "value": value, "unit": "path-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["config"], "sample_paths_per_step": n_paths},
        "cpu_baseline": {"value": value, "unit": "path-steps/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def measure(pcf, dist, job, name, steps, warmup, N=None):
    """K timed steps of one workload. Returns dict(device_s, wall_s, launches, units, price, se)."""
    import torch
    for i in range(warmup):
        run_ours_once(pcf, name, SEED0 + 1000 + i, N)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dist.barrier(job)
    t0 = time.perf_counter()
    dev, launches, last, abi = [], 0, None, 0.0
    for i in range(steps):
        last = run_ours_once(pcf, name, SEED0 + i, N)   # synchronous: returns after the result is on the host
        dev.append(last.seconds_kernel)
        abi += last.seconds_total
        launches += last.launches
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dist.barrier(job)
    wall = time.perf_counter() - t0
    # max over ranks, per step for the device clock and once for the wall clock
    red = dist.reduce_scalars(job, dev + [wall, abi], "max")
    tot_launch = dist.reduce_scalars(job, [float(launches)], "sum")[0]
    return dict(device_s=sum(red[:-2]), wall_s=red[-2], abi_s=red[-1], launches=int(tot_launch), units=last.units,
                price=last.price, se=last.std_error, t0=t0)


PARITY_CASES = (
    # (label, method, args, kwargs) -- small sizes of every method; N not divisible by the rank count; both mc_amer rules
    ("mc_asia", "mc_asia", (100, 100, .05, .2, 1, 1_000_001, 252, "call"), {"seed": 31}),
    ("mc_eur", "mc_eur", (100, 100, .05, .2, 1, 3_000_001, "put"), {"seed": 31}),
    ("mc_eur_multi", "mc_eur_multi", (100, 100, .05, .2, 1, 500_001, "call", 16, 0.5), {"seed": 31}),
    ("mc_eur_multi_eigen", "mc_eur_multi", (100, 100, .05, .2, 1, 200_001, "call", 4, 1.0), {"seed": 31}),
    ("mc_amer_put", "mc_amer", (100, 100, .05, .2, 1, 1_000_002, 50, "put"), {"seed": 31}),
    ("mc_amer_call", "mc_amer", (100, 110, .02, .75, 1, 200_000, 20, "call"), {"seed": 31}),
    ("mc_amer_lsm", "mc_amer", (100, 100, .05, .2, 1, 400_000, 50, "put"), {"seed": 31, "lsm": True}),
    ("mc_amer_tiny", "mc_amer", (100, 100, .05, .2, 1, 2, 5, "put"), {"seed": 31}),   # fewer pairs than ranks
    ("binom_embar", "binom", (100, 100, .05, .2, 1, 1_000_001, "call"), {}),
)
PARITY_TOL = 1e-13


def multi_gpu_parity(pcf, dist, job):
    """Every method at a small size: rank 0 alone on one GPU, then the N ranks together with the NVLink peer mailboxes and
    again with ncclAllReduce. Same Philox streams (keyed by global index), so only the summation order differs: relative
    differences must stay below 1e-13. Runs outside the timed region; returns the dict for the JSON line (rank 0)."""
    def price_all():
        return {lab: getattr(pcf, meth)(*a, **kw).price for lab, meth, a, kw in PARITY_CASES}

    single = None
    if job.rank == 0:
        pcf.init_rank(0, 1, job.local_rank, None)
        single = price_all()
        pcf.shutdown()
    dist.barrier(job)
    dist.init_library(job)
    out = {"tolerance": PARITY_TOL, "ranks": job.world, "modes": {}}
    modes = [("peer_mailbox", True)] if pcf.peer_active() else []
    modes.append(("nccl_allreduce", False))
    ok = True
    for mode, peer in modes:
        try:
            pcf.peer_enable(peer)
            multi = price_all()
            if job.rank == 0:
                errs = {lab: abs(multi[lab] - single[lab]) / max(abs(single[lab]), 1e-300) if single[lab] != 0
                        else abs(multi[lab]) for lab in multi}
                out["modes"][mode] = errs
                ok = ok and all(e <= PARITY_TOL for e in errs.values())
        except Exception as ex:  # a failure here must show up in the line, not kill the bench
            out["modes"][mode] = {"error": str(ex)}
            ok = False
        dist.barrier(job)
    if len(modes) == 2:
        pcf.peer_enable(True)   # the timed region runs in the default mode
    out["ok"] = bool(ok)
    if job.rank == 0:
        out["single_gpu_prices"] = single
    return out


def ncu_traffic(kernel_names, all_kernels=False):
    """dram__bytes_read.sum + dram__bytes_write.sum from the committed ncu launch list of this very command
    (profiles/ncu_traffic.json, written by tools/summarize_launches.py); None when no capture is committed.
    Default: bytes per launch of the DOMINANT kernel among `kernel_names` (most DRAM bytes). all_kernels=True: bytes per
    STEP summed over every listed kernel (launches of a kernel per step = its launches / the dominant kernel's launches
    in the capture) -- what a workload made of several kernels really moves."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        ks = json.load(open(path))["kernels"]
    except (ValueError, KeyError):
        return None
    hits = []
    for frag in kernel_names.split("+"):
        for k, v in ks.items():
            if k.startswith(frag):
                hits.append(v)
    if not hits:
        return None
    dom = max(hits, key=lambda v: v["dram_bytes_per_launch"])
    if not all_kernels:
        return dom["dram_bytes_per_launch"]
    steps = max(min(v["launches"] for v in hits if v["dram_bytes_per_launch"] > 0.25 * dom["dram_bytes_per_launch"]), 1)
    return sum(v["dram_bytes_per_launch"] * v["launches"] for v in hits) / steps


_NCU_DIGEST = {"mc_asia": "mc_asia", "mc_eur": "mc_eur", "mc_eur_stated": "mc_eur", "mc_eur_multi": "basket_equi",
               "mc_basket_general": "basket_general", "mc_amer": "amer_sweep", "binom_embar": "binom_screen",
               "binom_embar_max": "binom_screen", "binom_embar_noscreen": "binom_noscreen", "binom_vanilla_amer": "tree_amer"}


def ncu_counters(name):
    """The counters BASELINE.json's north_star names, from the committed `ncu --set full` capture of the workload's
    dominant kernel as shipped (profiles/r2g_ncu_<kernel>.txt, digested by tools/ncu_summary.py): FP64-pipe and
    issue-slot utilisation, DRAM throughput. Static evidence (a capture, not this run); None when there is no digest."""
    key = _NCU_DIGEST.get(name)
    path = os.path.join(ROOT, "profiles", f"r2g_ncu_{key}.txt") if key else None
    if not path or not os.path.exists(path):
        return None
    want = {"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_pct",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct"}
    out = {}
    for line in open(path):
        parts = line.split()
        if len(parts) >= 4 and parts[0] in want and want[parts[0]] not in out:
            try:
                out[want[parts[0]]] = round(float(parts[-1]), 2)
            except ValueError:
                pass
    if out:
        out["source"] = f"profiles/r2g_ncu_{key}.txt"
    return out or None


def binom_executed_slots(N):
    """FP64 instructions per term the binomial kernel EXECUTES with its screening pass on: every pair costs
    the two-logarithm screen (32), and only pairs with a term whose log-weight is above the -712 underflow rule go on to the
    full saddle-point routine (150 more). Live pairs from the Gaussian envelope of the weights around N p."""
    import math
    _, _, p, q = __import__("oracle").binom_params(P["r"], P["sigma"], P["T"], N)
    npq = N * p * q
    half_width = math.sqrt(2.0 * npq * max(712.0 - 0.5 * math.log(2 * math.pi * npq), 1.0))
    pairs = (N + 1) // 2
    live = min(pairs, half_width + abs(N * p - N / 2.0) + 1)
    return (32.0 * pairs + 150.0 * live) / (N + 1), live / pairs


def roofline_for(name, units_per_s, fp64_dfma_per_s, hbm_bytes_per_s, hbm_src):
    r = _roofline_for(name, units_per_s, fp64_dfma_per_s, hbm_bytes_per_s, hbm_src)
    w = WORKLOADS[name]
    if "exec_cycles" in w:
        # The algorithmic slot count (libdevice-based) is above what this build executes, so `frac` can exceed
        # 1. This is the same rate against a MODEL of the work the kernel issues: an FP64 instruction holds the FP64 pipe
        # for 2 cycles per warp, a Philox IMAD.WIDE holds the fmaheavy pipe for 4 (ncu counters of tools/ubench,
        # profiles/r2g_ncu_ubench_pipes.txt), and a loop that mixes the two runs at 0.85-0.87 of the SUM of both, far
        # from their maximum (same file: 137 us DFMA alone, 287 us IMAD.WIDE alone, 369 us mixed), so the budget per unit
        # is taken as 2 #FP64 + 4 #IMAD.WIDE cycles. It is a model, not a counter: the ncu counters of the shipped kernel
        # are in `ncu` below.
        r["executed_pipe_frac"] = units_per_s * w["exec_cycles"] / (2.0 * fp64_dfma_per_s)
        r["executed_work"] = (f"{w['exec_cycles']:g} cycles per warp and unit in the additive FP64 + IMAD.WIDE model "
                              f"({w['exec_src']})")
    nc = ncu_counters(name)
    if nc:
        r["ncu"] = nc
    if name.startswith("binom_embar") and name != "binom_embar_noscreen":
        # the screening pass settles most pairs with ~32 FP64 instructions, so the ALGORITHMIC 85 slots per term (which the
        # unscreened entry is quoted on) overstate what this run executed: report the rate against executed work and mark
        # the algorithmic fraction as not creditable (shortcuts do not change the per-unit figure)
        try:
            slots, live = binom_executed_slots(WORKLOADS[name]["N"])
            r["executed_frac"] = units_per_s * slots / fp64_dfma_per_s
            r["executed_work"] = (f"{slots:.1f} FP64 instructions per term: 32 per pair for the screen + 150 more for the "
                                  f"{100 * live:.2f} % of pairs that reach the full routine (DESIGN 4.4)")
            r["frac_note"] = ("`frac` counts the algorithmic 85 slots for every term although the screen settles most of "
                              "them early; the creditable figures are `executed_frac` here and `frac` of the screening-OFF "
                              "entry")
        except Exception as ex:  # the oracle's lattice helper is missing: leave the algorithmic figure alone
            r["executed_work"] = f"unavailable ({ex})"
    r["traffic"] = ncu_traffic(w["kernel"], w.get("traffic_all", False))
    if r["traffic"] is not None:
        if w.get("traffic_all"):
            r["traffic_source"] = ("profiles/ncu_traffic.json: DRAM read+write bytes per STEP summed over "
                                   + w["kernel"].replace("+", ", ") + " (ncu)")
            if w["bound"] == "hbm":
                # executed view: the bytes the step really moved / its time, against the same peak
                t_step = WORKLOADS[name]["N"] * w["steps_per_unit"] / units_per_s
                r["executed_achieved"] = r["traffic"] / t_step / 1e9
                r["executed_frac"] = r["executed_achieved"] / r["peak"]
        else:
            r["traffic_source"] = "profiles/ncu_traffic.json: DRAM read+write bytes per launch of the dominant kernel (ncu)"
    return r


def _roofline_for(name, units_per_s, fp64_dfma_per_s, hbm_bytes_per_s, hbm_src):
    w = WORKLOADS[name]
    if w["bound"] == "fp64":
        ach = units_per_s * w["slots"] * 2 / 1e12
        peak = fp64_dfma_per_s * 2 / 1e12
        return {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "kernel": w["kernel"], "algorithmic": f"{w['slots']:g} FP64 issue slots per unit ({w.get('algo_src', 'SURVEY 8d')}), FMA = 2 flop",
                "peak_source": "measured in this run: DFMA-chain microbenchmark (pcf_fp64_peak); MEASURED_PEAKS.json has no FP64 figure"}
    ach = units_per_s * w["bytes"] / 1e9
    peak = hbm_bytes_per_s / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "kernel": w["kernel"], "algorithmic": f"{w['bytes']:g} B per path-step (SURVEY 8d: 8 B path write + 28 B sweep)",
            "peak_source": hbm_src}


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line. Libraries loaded later (NCCL's version banner, torch.distributed) write to
    file descriptor 1 behind Python's back, so fd 1 is pointed at stderr for the rest of the process and the JSON line
    goes to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mc_asia", choices=sorted(WORKLOADS))
    ap.add_argument("--paths", type=int, default=0, help="override the workload's path count (debug)")
    ap.add_argument("--ref-paths", type=int, default=2_000_000, help="--impl reference: paths per step (bounded sample)")
    ap.add_argument("--cpu-paths", type=int, default=12_000_000, help="cpu_baseline sample (paths)")
    ap.add_argument("--no-others", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    if args.warmup < 3:
        args.warmup = 3

    from parcompfin_b200 import dist
    rank, world, local = dist.env_job()
    if args.impl == "reference":
        bench_reference(args, rank)
        return

    import torch
    import parcompfin_b200 as pcf
    pcf.load_library()  # fails loudly when the CUDA extension is missing: there is no fallback
    job = dist.setup()
    parity = None
    if job.world > 1:
        parity = multi_gpu_parity(pcf, dist, job)   # outside the timed region; leaves the library up in its default mode
    else:
        pcf.init(args.gpus)  # single process: N GPUs driven in-process (N = 1 in the default run)
    n_gpus = pcf.world_size()

    name = args.workload
    w = WORKLOADS[name]
    sampler = ClockSampler(local) if job.rank == 0 else None
    if sampler:
        sampler.start()
    t_wall0 = time.time()
    m = measure(pcf, dist, job, name, args.steps, args.warmup, args.paths or None)
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    total_units = m["units"] * args.steps
    value = total_units / m["device_s"]
    e2e = total_units / m["wall_s"]

    # roofline denominators, measured on this GPU right after the timed region
    fp64_peak = pcf.fp64_peak(0.25)
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        hbm = json.load(open(peaks_file))["hbm_gbs"] * 1e9
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        hbm, hbm_src = 6.65e12, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"

    others = []
    if not args.no_others:
        for other in WORKLOADS:
            if other == name:
                continue
            try:
                wo = WORKLOADS[other]
                k = 30 if wo.get("latency") else 2
                mo = measure(pcf, dist, job, other, k, 3)
                ups = mo["units"] * k / mo["device_s"]
                others.append({"workload": WORKLOADS[other]["config"], "value": ups, "unit": WORKLOADS[other]["unit"],
                               "e2e": mo["units"] * k / mo["wall_s"], "ms_per_step": 1e3 * mo["device_s"] / k,
                               "price": mo["price"], "std_error": mo["se"], "gpu_launches": mo["launches"],
                               # a path that does not shard (the tree) runs as replicas: its rate is per GPU already
                               "roofline": roofline_for(other, ups if WORKLOADS[other].get("replicas") else ups / n_gpus,
                                                        fp64_peak, hbm, hbm_src)})
                if wo.get("latency"):
                    # launch-latency-bound sizes: microseconds per call, and the same for a ONE-unit call of the method
                    # (the floor: launch, grid reduction, completion, result read-back and nothing else)
                    fl = measure(pcf, dist, job, other, k, 3, N=2)
                    others[-1]["latency_us"] = {
                        "device": 1e6 * mo["device_s"] / k, "c_abi_call": 1e6 * mo["abi_s"] / k,
                        "python_call": 1e6 * mo["wall_s"] / k,
                        "floor_device": 1e6 * fl["device_s"] / k, "floor_c_abi_call": 1e6 * fl["abi_s"] / k,
                        "note": "device = CUDA events around the kernel; c_abi_call = wall clock of the pcf_* call itself "
                                "(parameter struct in, result out, one stream synchronise); python_call adds the ctypes "
                                "binding; floor = the same call on 2 units"}
                if "assets" in WORKLOADS[other]:  # the basket is also quoted in asset-steps
                    others[-1]["asset_steps_per_sec"] = ups * WORKLOADS[other]["assets"]
                if job.rank == 0 and n_gpus == 1 and not args.no_cpu:
                    others[-1]["cpu_baseline"] = cpu_reference_other(other, host_threads())
            except Exception as ex:  # e.g. the path store does not fit
                others.append({"workload": WORKLOADS[other]["config"], "error": str(ex)})

    cpu = None
    if job.rank == 0 and n_gpus == 1 and not args.no_cpu:
        threads = host_threads()
        v, kind, sec = cpu_reference_asia(args.cpu_paths, threads)
        cpu = {"value": v, "unit": "path-steps/s", "cores": threads, "kind": kind,
               "sample": f"{args.cpu_paths} of 1e9 paths x 252 dates, {sec:.1f} s, mc_asia_omp "
                         + ("(unmodified reference source, -O2)" if kind == "reference" else "(oracle port)")}
        try:  # the reference ships without -O (Makefile:9); reported next to the -O2 headline
            import oracle
            if kind == "reference" and os.path.exists(os.path.join(oracle.REF_DIR, "mc_asia_omp_O0")):
                n0 = max(args.cpu_paths // 4, 1000)
                row = oracle.ref_row("mc_asia_omp_O0", "call", 100, 100, 0.05, 0.2, 1, n0, 252, threads)
                cpu["as_shipped_no_O"] = {"value": n0 * 252 / float(row[12]), "unit": "path-steps/s",
                                          "sample": f"{n0} paths x 252 dates, {float(row[12]):.1f} s, mc_asia_omp built with "
                                                    "the reference's own flags (no -O, Makefile:9)"}
        except Exception as ex:
            cpu["as_shipped_no_O"] = {"error": str(ex)}

    if job.rank == 0:
        line = {
            "metric": "path_steps_per_sec", "value": value, "unit": w["unit"], "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * m["device_s"] / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["config"], "paths": args.paths or w["N"], "dates": w["M"],
                       "parallelism": f"paths sharded by global index over {n_gpus} GPU(s)",
                       "collective": ("none (1 GPU)" if n_gpus == 1 else
                                      "NVLink peer-memory mailbox: P2P stores issued by the reducing kernel (csrc/xchg.cuh)"
                                      if pcf.peer_active() else "ncclAllReduce on the compute stream"),
                       "l2": "no input arrays: the kernel reads no global memory (parameters only), nothing to flush; "
                             "the mc_amer entry in `others` streams a 40 GB path store (>> 126 MB L2)",
                       "price": m["price"], "std_error": m["se"]},
            "e2e": {"value": e2e, "unit": w["unit"], "h2d_bytes_per_step": 104 * n_gpus, "d2h_bytes_per_step": 20 * n_gpus},
            "gpu_launches": m["launches"],
            "roofline": roofline_for(name, value / n_gpus, fp64_peak, hbm, hbm_src),
            "cpu_baseline": cpu,
            "clocks": clocks,
            "others": others,
        }
        if parity is not None:
            line["parity"] = parity
        emit(line)
    dist.barrier(job)
    pcf.shutdown()
    dist.teardown(job)


if __name__ == "__main__":
    main()
