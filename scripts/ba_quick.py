"""C3 / C4 bundle adjustment on one GPU: lambda-trials/s and per-phase device times (no CPU leg)."""
import sys
sys.path.insert(0, '/root/repo')
from ptam_cg_b200.bench_ba import bench_ba
from ptam_cg_b200.capi import product_lib
prod = product_lib()
for cfg in sys.argv[1:] or ("C3", "C4"):
    o = bench_ba(prod, config=cfg, reps=3)
    print(cfg, round(o["value"], 1), "trials/s", round(o["compute_ms"], 2), "ms", o["lambda_trials"], "trials",
          {k: round(v, 4) for k, v in o["phases_ms_per_call"].items()}, flush=True)
