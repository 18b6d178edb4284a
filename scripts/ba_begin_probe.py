"""Times ptam_bundle_begin (host CSR build + arena + upload) apart from the LM loop at C3 / C4."""
import sys, time
sys.path.insert(0, '/root/repo')
from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Bundle, product_lib
from ptam_cg_b200.bench_ba import CONFIGS
prod = product_lib()
for cfg in ("C3", "C4"):
    g = synth.make_ba_graph(**CONFIGS[cfg])
    for rep in range(3):
        b = Bundle(prod, g["width"], g["height"])
        t0 = time.perf_counter(); b.add_graph(g); t1 = time.perf_counter()
        b.begin(); b.synchronize(); t2 = time.perf_counter()
        acc = b.Compute(); b.synchronize(); t3 = time.perf_counter()
        s = b.stats()
        print(cfg, rep, f"add_graph {1e3*(t1-t0):.2f} ms  begin {1e3*(t2-t1):.2f} ms  compute {1e3*(t3-t2):.2f} ms  trials {s.lambda_trials}  per-trial {1e3*(t3-t2)/s.lambda_trials:.3f} ms", flush=True)
        b.close()
