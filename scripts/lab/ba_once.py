"""One Bundle::Compute on a BASELINE config (C3 / C4) with the library's debug timers on stderr."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from ptam_cg_b200.capi import product_lib
from ptam_cg_b200 import bench_ba, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
g = synth.make_ba_graph(**bench_ba.CONFIGS[cfg])
lib = product_lib()
for r in range(3):
    b = bench_ba._run(lib, g, 0, None, profile=False)
    b.synchronize()
    t0 = time.perf_counter(); acc = b.Compute(); b.synchronize(); dt = time.perf_counter() - t0
    s = b.stats()
    print(cfg, "run", r, "wall %.2f ms" % (dt * 1e3), "accepted", acc, "trials", s.lambda_trials, "outliers", s.n_outliers, flush=True)
    b.close()
