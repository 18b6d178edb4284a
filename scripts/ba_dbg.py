import sys, ctypes
sys.path.insert(0, '/root/repo')
from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Bundle, product_lib
from ptam_cg_b200.bench_ba import CONFIGS
g = synth.make_ba_graph(**CONFIGS["C4"])
lib = product_lib()
b = Bundle(lib, g["width"], g["height"], max_iterations=2)
b.add_graph(g)
print(b.Compute())
out = (ctypes.c_longlong * 8)()
lib.cdll.ptam_debug_read(out)
print([x / (2 * 47) for x in out], "cycles per panel: load, ldlt64, writeback, rowload, rowsolve, rowstore")
