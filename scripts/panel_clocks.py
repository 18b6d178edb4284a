"""Debug probe (library built with -DPTAM_PANEL_DEBUG): cycles CTA 0 of k_ldlt_panel spends per phase."""
import sys, ctypes as C
sys.path.insert(0, '/root/repo')
from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Bundle, product_lib
from ptam_cg_b200.bench_ba import CONFIGS
prod = product_lib()
names = ["load", "pending update", "factor", "store diag", "load rows", "row solve", "store W/L"]
for cfg in ("C3", "C4"):
    g = synth.make_ba_graph(**CONFIGS[cfg])
    b = Bundle(prod, g["width"], g["height"]); b.add_graph(g)
    z = (C.c_longlong * 8)()
    prod.cdll.ptam_debug_read(z); base = list(z)
    b.Compute(); s = b.stats()
    prod.cdll.ptam_debug_read(z)
    n = 6 * (len(g["cam_fixed"]) - 1)
    panels = ((n + 63) // 64) * s.lambda_trials
    print(cfg, "panels", panels, {nm: round((z[i] - base[i]) / panels / 1.9e3, 2) for i, nm in enumerate(names)}, "us per panel launch (1.9 GHz)")
