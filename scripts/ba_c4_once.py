import sys
sys.path.insert(0, '/root/repo')
from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Bundle, product_lib
from ptam_cg_b200.bench_ba import CONFIGS
cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
g = synth.make_ba_graph(**CONFIGS[cfg])
b = Bundle(product_lib(), g["width"], g["height"], max_iterations=int(sys.argv[2]) if len(sys.argv) > 2 else 2)
b.add_graph(g)
print(b.Compute(), b.stats().lambda_trials)
