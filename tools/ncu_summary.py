#!/usr/bin/env python
"""Text summary of an .ncu-rep for profiles/: `python tools/ncu_summary.py <report.ncu-rep> ["header line"]`.
Prints, per captured launch, the metrics the DESIGN/bench numbers are argued from (durations, DRAM bytes, pipe and
issue utilisation, occupancy limits, warp stall reasons above 0.1 per issued instruction)."""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct")


def main(rep, header=None):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"]).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    names, units = rows[0], rows[1]
    if header:
        print(header + "\n")
    for r in rows[2:]:
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        print("== " + d.get("Kernel Name", "?"))
        for k in KEEP:
            if k in d and d[k] != "":
                print(f"  {k} [{u[k]}] = {d[k]}")
        for k in sorted(d):
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(d[k]) >= 0.1:
                        print(f"  {k} [{u[k]}] = {d[k]}")
                except ValueError:
                    pass
        print()


if __name__ == "__main__":
    main(*sys.argv[1:3])
