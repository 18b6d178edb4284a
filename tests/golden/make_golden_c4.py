"""Generates tests/golden/bundle_c4_reference.npz: the reference's OWN Bundle::Compute (src/Bundle.cc of
/root/reference, compiled in place into oracle/_ref/libref_ptam.so by oracle/Makefile.ref) run to completion on
BASELINE config C4 (500 keyframes x 100 000 points x 600 000 measurements, synth.make_ba_graph seed 43), and the
oracle (the restatement with the specified atan, the CUDA product's numeric contract) on the same graph.
Run from the repo root, here where /root/reference exists:  python tests/golden/make_golden_c4.py   (a few CPU-minutes
per library).  The graph itself is not stored (it is regenerated from its seed); stored are the integer outcomes in
full and the states as every 25th point + all cameras, plus float64 sums as a whole-array check.
Also stores the per-LM-step trace (error, sigma^2, outliers so far) of both, which is what the -m gpu test compares
step by step."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from ptam_cg_b200 import synth  # noqa: E402
from ptam_cg_b200.bench_ba import CONFIGS  # noqa: E402
from ptam_cg_b200.capi import Bundle  # noqa: E402
from oracle.binding import oracle_lib, ref_lib  # noqa: E402

STRIDE = 25


def run(lib, g, tag, stepwise=True):
    b = Bundle(lib, g["width"], g["height"])
    b.add_graph(g)
    t0 = time.perf_counter()
    trace = []
    if stepwise:
        b.begin()
    else:  # the reference's class has no step-wise entry (Bundle.h:110-118): one Compute()
        b.Compute()
    while True:
        if stepwise:
            b.lm_step()
        s = b.stats()
        trace.append((s.lambda_trials, s.accepted, s.n_outliers, s.sigma_squared, s.last_error, s.last_new_error, s.lambda_))
        print(f"{tag}: step {len(trace)} trials {s.lambda_trials} accepted {s.accepted} outliers {s.n_outliers} "
              f"err {s.last_error:.6f} -> {s.last_new_error:.6f} ({time.perf_counter() - t0:.0f} s)", flush=True)
        if s.converged or s.hit_max_iterations or not stepwise:
            break
    pts, cams = b.get_points(), b.get_cameras()
    out = {f"{tag}_accepted": s.accepted, f"{tag}_lambda_trials": s.lambda_trials, f"{tag}_lm_steps": s.lm_steps,
           f"{tag}_converged": s.converged, f"{tag}_outliers": b.GetOutlierMeasurements().astype(np.int32),
           f"{tag}_cameras": cams, f"{tag}_points_sub": pts[::STRIDE].copy(), f"{tag}_points_sum": pts.sum(0),
           f"{tag}_points_abs_sum": np.abs(pts).sum(0), f"{tag}_trace": np.array(trace, np.float64),
           f"{tag}_seconds": time.perf_counter() - t0}
    b.close()
    return out


def main():
    g = synth.make_ba_graph(**CONFIGS["C4"])
    out = dict(stride=STRIDE, config=np.array([CONFIGS["C4"][k] for k in ("n_cams", "n_points", "n_meas", "seed")]))
    ref = ref_lib()
    assert ref is not None, "oracle/_ref is not built (needs /root/reference)"
    out.update(run(ref, g, "ref", stepwise=False))
    out.update(run(oracle_lib(), g, "orc"))
    np.savez_compressed(ROOT / "tests/golden/bundle_c4_reference.npz", **out)
    print("written", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
