"""BA leg of bench.py: Bundle::Compute on the synthetic C3 graph (50 keyframes x 5000 points x
20000 measurements), lambda-trials/s on one B200 next to the CPU oracle port on one core."""
import time

from . import synth
from .capi import Bundle


def bench_ba(prod, device=0, n_cams=50, n_points=5000, n_meas=20000, seed=42, reps=5, cpu_lib=None):
    """cpu_lib: an already-loaded CPU library exporting the same ABI (bench.py passes the oracle for the
    cpu_baseline leg); this package never loads it itself."""
    g = synth.make_ba_graph(n_cams, n_points, n_meas, seed=seed)
    best = None
    for r in range(reps + 1):  # first repetition is the warm-up
        b = Bundle(prod, g["width"], g["height"], device=device)
        b.add_graph(g)
        b.synchronize()
        l0 = b.launch_count()
        t0 = time.perf_counter()
        acc = b.Compute()
        dt = time.perf_counter() - t0
        s = b.stats()
        if r and (best is None or dt < best[0]):
            best = (dt, acc, s.lambda_trials, s.lm_steps, s.n_outliers, b.launch_count() - l0)
        b.close()
    dt, acc, trials, steps, outl, launches = best
    out = {"workload": f"C3: Bundle::Compute LM, {n_cams} keyframes x {n_points} points x {n_meas} measurements",
           "value": trials / dt, "unit": "lambda-trials/s", "accepted_steps_per_s": acc / dt, "compute_ms": dt * 1e3,
           "lambda_trials": trials, "accepted": acc, "lm_steps": steps, "outliers": outl, "gpu_launches": launches,
           "timing": "wall clock around ptam_bundle_compute (host LM control + device phases), best of %d" % reps}
    if cpu_lib is not None:
        o = Bundle(cpu_lib, g["width"], g["height"])
        o.add_graph(g)
        t0 = time.perf_counter()
        acc_o = o.Compute()
        dto = time.perf_counter() - t0
        so = o.stats()
        out["cpu_baseline"] = {"value": so.lambda_trials / dto, "unit": "lambda-trials/s", "cores": 1, "kind": "port",
                               "sample": f"one full Compute() on the same graph: {so.lambda_trials} trials, {acc_o} accepted, {dto * 1e3:.0f} ms"}
    return out
