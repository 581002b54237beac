"""Worker for tests/test_multi_gpu.py: one rank of a torchrun job; rank 0 prints one JSON line of prices."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parcompfin_b200 as pcf
from parcompfin_b200 import dist

job = dist.setup()
dist.init_library(job)
P1 = (100, 100, 0.05, 0.2, 1)
out = {"world": pcf.world_size(), "peer": pcf.peer_active()}
out["asia"] = pcf.mc_asia(*P1, 1_000_001, 252, "call", seed=31).price
out["eur"] = pcf.mc_eur(*P1, 3_000_001, "put", seed=31).price
out["basket"] = pcf.mc_eur_multi(*P1, 500_001, "call", 16, 0.5, seed=31).price
out["amer"] = pcf.mc_amer(*P1, 1_000_002, 50, "put", seed=31).price
out["amer_call"] = pcf.mc_amer(100, 110, 0.02, 0.75, 1, 200_000, 20, "call", seed=31).price
out["binom"] = pcf.binom(*P1, 1_000_001, "call").price
out["amer_repeat"] = pcf.mc_amer(*P1, 1_000_002, 50, "put", seed=31).price
# fewer units than ranks: trailing ranks hold an EMPTY shard and must still take part in every exchange
out["amer_tiny"] = pcf.mc_amer(*P1, 2, 5, "put", seed=31).price
out["asia_tiny"] = pcf.mc_asia(*P1, 1, 7, "call", seed=31).price
out["eur_tiny"] = pcf.mc_eur(*P1, 1, "put", seed=31).price
out["binom_tiny"] = pcf.binom(*P1, 1, "call").price
out["amer_lsm"] = pcf.mc_amer(*P1, 400_000, 50, "put", seed=31, lsm=True).price
if job.rank == 0:
    print("MGPU " + json.dumps(out), flush=True)
dist.barrier(job)
pcf.shutdown()
dist.teardown(job)
