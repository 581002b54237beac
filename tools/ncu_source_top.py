#!/usr/bin/env python
"""Source-page digest of an .ncu-rep captured with --import-source on: `python tools/ncu_source_top.py <report> [N]`.
Prints the share of every warp-stall reason over all samples of the kernel and the N instructions (SASS) with the most
samples, each with its two dominant stall reasons -- the evidence behind "what this kernel waits for" in DESIGN.md."""
import csv
import io
import subprocess
import sys


def main(rep, top=25):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                                  stderr=subprocess.DEVNULL).decode()
    blocks, cur = [], []
    for line in raw.splitlines():
        if line.startswith('"Kernel Name"'):
            if cur:
                blocks.append(cur)
            cur = [line]
        elif cur:
            cur.append(line)
    if cur:
        blocks.append(cur)
    for b in blocks:
        name = next(csv.reader([b[0]]))[1]
        rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
        hdr, data = rows[0], [r for r in rows[1:] if len(r) == len(rows[0])]
        ix = {h: i for i, h in enumerate(hdr)}
        if "# Samples" not in ix:
            continue
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        print("== " + name[:150])
        print(f"   {len(data)} SASS instructions, {tot} warp samples, "
              f"{sum(int(r[ix['Instructions Executed']]) for r in data)} warp instructions executed")
        agg = sorted(((sum(int(r[ix[c]]) for r in data), c) for c in cols), reverse=True)
        print("   stall reasons: " + ", ".join(f"{c[6:]} {100 * v / tot:.1f}%" for v, c in agg if v * 200 >= tot))
        base = int(data[0][0], 16)
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
            st = sorted(((int(r[ix[c]]), c[6:]) for c in cols), reverse=True)[:2]
            print(f"   +{int(r[0], 16) - base:05x} {100 * int(r[ix['# Samples']]) / tot:5.1f}%  x{r[ix['Instructions Executed']]:>10}  "
                  f"{r[1].strip()[:58]:58s} {st[0][1]} {st[0][0]}, {st[1][1]} {st[1][0]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
