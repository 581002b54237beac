#!/usr/bin/env python
"""Turns the ncu launch list of a bench.py run into the two artefacts bench.py and the judge read:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 \\
        --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu
    python tools/summarize_launches.py gpurun_out/launches.csv profiles/rNN_launch_list_summary.txt profiles/ncu_traffic.json

Per-launch times under ncu are cold-cache and serialised: the SHARES are what must agree with the live CUDA-event
timing, not the absolute values. The JSON holds dram read+write bytes per launch per kernel (bench.py's
`roofline.traffic`)."""
import collections
import csv
import json
import re
import sys


def main(src, txt_out, json_out):
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault(r["ID"], {"kernel": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        name = r["Metric Name"]
        if name == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[unit]
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
        per[r["ID"]][name] = v
    agg = collections.OrderedDict()
    for d in per.values():
        short = re.sub(r"\(.*", "", d["kernel"]).replace("void ", "").replace("pcf::", "")
        a = agg.setdefault(short, {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["ms"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    total = sum(a["ms"] for a in agg.values()) or 1.0
    out = []
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        out.append(f"{k[:52]:52s} n={a['n']:4d} total_ms={a['ms']:10.3f} avg_ms={a['ms'] / a['n']:9.4f} "
                   f"share={a['ms'] / total:.3f} avg_dram_rd_GB={a['rd'] / a['n'] / 1e9:.3f} "
                   f"avg_dram_wr_GB={a['wr'] / a['n'] / 1e9:.3f}")
    with open(txt_out, "w") as f:
        f.write("\n".join(out) + "\n")
    traffic = {k: {"launches": a["n"], "dram_bytes_per_launch": (a["rd"] + a["wr"]) / a["n"],
                   "avg_ms_under_ncu": a["ms"] / a["n"]} for k, a in agg.items()}
    with open(json_out, "w") as f:
        json.dump({"source": f"{txt_out} (ncu launch list of `python bench.py`, dram__bytes_read.sum + dram__bytes_write.sum "
                             "per launch, averaged over the launches of each kernel)", "kernels": traffic}, f, indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main(*sys.argv[1:4])
