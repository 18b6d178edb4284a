import sys, json, time
sys.path.insert(0, '/root/repo')
from ptam_cg_b200.capi import product_lib
from ptam_cg_b200.bench_ba import bench_ba
t=time.time()
for cfg in ("C3","C4"):
    out = bench_ba(product_lib(), 0, cfg, reps=2)
    print(cfg, json.dumps(out), flush=True)
print('total', time.time()-t)
