"""Dev tool: T(N) of every pinned CTA shape on a grid of N; the slope between two grid points is the cost of one
layer at that width, from which tree_pick_shape's table is derived (tools/tree_shape_table.py)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
pcf.init(1)
P = (100., 100., .05, .2, 1.)
shapes = sys.argv[1:] or ["1204", "1304", "1404", "1604", "1108", "1208", "1308", "1408", "1608", "1212", "1312", "1412", "1612",
                          "1116", "1216", "1316", "1416", "1616", "1120", "1808", "1812", "1816", "1820",
                          "14044", "13084", "12084", "12124"]
grid = [2000, 5000, 10000, 20000, 30000, 40000, 50000, 60000, 70000, 80000, 90000, 100000, 125000, 150000, 200000,
        250000, 300000, 400000, 500000, 700000, 1000000]
out = {}
for shape in shapes:
    os.environ["PCF_TREE"] = shape
    for fn, name in ((pcf.binom_vanilla_eur, "eur"), (pcf.binom_vanilla_amer, "amer")):
        ts = []
        for N in grid:
            if N > 400000 and shape[1] in "12" and shape[2:4] in ("04", "08"):
                ts.append(None)  # small shapes at huge N: not candidates, skip the time
                continue
            best = min(fn(*P, N, "put").seconds_kernel for _ in range(2 if N >= 200000 else 3))
            ts.append(best)
        out[f"{shape}:{name}"] = ts
        print(shape, name, " ".join("-" if t is None else f"{t*1e3:.3f}" for t in ts), flush=True)
path = os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "tree_shape_times.json")
if sys.argv[1:] and os.path.exists(path):  # a partial run extends the table
    old = json.load(open(path))
    if old.get("grid") == grid:
        old["seconds"].update(out)
        out = old["seconds"]
json.dump({"grid": grid, "seconds": out}, open(path, "w"))
pcf.shutdown()
