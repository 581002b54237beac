#!/usr/bin/env python
"""Pipe-level counters of every launch in an .ncu-rep: `python tools/ncu_pipes.py <report>`. One line per launch with the
warp instructions executed per pipe (sm__inst_executed_pipe_*), the cycles each pipe was busy and the issue-slot use --
the ncu-side evidence for which pipe an instruction class (e.g. IMAD.WIDE) occupies."""
import csv
import io
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main(rep):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    names, units = rows[0], rows[1]
    want = [n for n in names if n.startswith(("sm__inst_executed_pipe_", "sm__pipe_")) and
            n.endswith((".sum", ".avg.pct_of_peak_sustained_active", ".avg", ".max"))]
    for r in rows[2:]:
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        print("== pipes: " + d.get("Kernel Name", "?")[:120])
        for k in ("sm__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_active.sum", "smsp__issue_active.sum",
                  "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__time_duration.sum"):
            if k in d and d[k] != "":
                print(f"  {k} [{u[k]}] = {d[k]}")
        for k in sorted(want):
            v = num(d.get(k, ""))
            if v:
                print(f"  {k} [{u[k]}] = {d[k]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
