"""GPU-side check of path B: determinism (bit-identical repeat runs), parity with the oracle and phase times.
usage: python scripts/ba_check.py [C3|C4|small] [--oracle]   (test / measurement tool, not product)"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ptam_cg_b200 import synth  # noqa: E402
from ptam_cg_b200.capi import Bundle, product_lib  # noqa: E402
from ptam_cg_b200.bench_ba import CONFIGS  # noqa: E402


def run(lib, g, profile=False, **kw):
    b = Bundle(lib, g["width"], g["height"], **kw)
    b.add_graph(g)
    if profile:
        b.set_profiling(True)
    t0 = time.perf_counter()
    acc = b.Compute()
    b.synchronize() if hasattr(b, "synchronize") else None
    dt = time.perf_counter() - t0
    s = b.stats()
    out = dict(acc=acc, trials=s.lambda_trials, steps=s.lm_steps, outliers=b.GetOutlierMeasurements().copy(), pts=b.get_points(), cams=b.get_cameras(),
               err=s.last_error, sigma2=s.sigma_squared, wall=dt, phases=b.phase_times() if profile else None)
    b.close()
    return out


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "C3"
    cfg = CONFIGS.get(which, dict(n_cams=10, n_points=500, n_meas=2500, seed=9))
    g = synth.make_ba_graph(**cfg)
    prod = product_lib()
    a = run(prod, g)
    b = run(prod, g)
    c = run(prod, g, profile=True)
    same = (a["acc"] == b["acc"] and a["trials"] == b["trials"] and np.array_equal(a["outliers"], b["outliers"])
            and np.array_equal(a["pts"], b["pts"]) and np.array_equal(a["cams"], b["cams"]) and a["err"] == b["err"])
    print(f"{which}: accepted {a['acc']} trials {a['trials']} steps {a['steps']} outliers {len(a['outliers'])} wall {b['wall']*1e3:.1f} ms; repeat run bit-identical: {same}")
    print("phases ms/call:", {k: round(v[0] / v[1], 4) if v[1] else None for k, v in c["phases"].items()})
    # step 0: reduced system vs oracle, bit for bit
    if "--oracle" in sys.argv:
        from oracle.binding import oracle_lib
        orc = oracle_lib()
        n = 6 * int((np.asarray(g["cam_fixed"]) == 0).sum())
        o, p = Bundle(orc, g["width"], g["height"]), Bundle(prod, g["width"], g["height"])
        o.add_graph(g); p.add_graph(g)
        o.begin(); p.begin()
        o.lm_step(); p.lm_step()
        So, eo = o.reduced_system(n)
        Sp, ep = p.reduced_system(n)
        print("step 0: S bit-equal", np.array_equal(So, Sp), "vE bit-equal", np.array_equal(eo, ep),
              "max |dS|/max|S|", np.abs(So - Sp).max() / np.abs(So).max(), "sigma2 equal", o.stats().sigma_squared == p.stats().sigma_squared,
              "err", o.stats().last_error, p.stats().last_error, "new", o.stats().last_new_error, p.stats().last_new_error)
        print("step 0: max |d pts|", np.abs(o.get_points() - p.get_points()).max(), "max |d cams|", np.abs(o.get_cameras() - p.get_cameras()).max())
        t0 = time.perf_counter()
        r = run(orc, g)
        print(f"oracle whole run {time.perf_counter() - t0:.1f} s: accepted {r['acc']} trials {r['trials']} outliers {len(r['outliers'])}")
        print("whole run: accepted/trials equal", r["acc"] == a["acc"], r["trials"] == a["trials"], "outliers equal", np.array_equal(r["outliers"], a["outliers"]),
              "max |d pts|", np.abs(r["pts"] - a["pts"]).max(), "max |d cams|", np.abs(r["cams"] - a["cams"]).max(), "err", r["err"], a["err"])


if __name__ == "__main__":
    main()
