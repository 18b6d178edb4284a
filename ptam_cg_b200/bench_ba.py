"""BA legs of bench.py: Bundle::Compute on the synthetic graphs of SURVEY.md §8d.

C3: 50 keyframes x 5 000 points x 20 000 measurements on one B200.
C4: 500 keyframes x 100 000 points x 600 000 measurements, one B200 or sharded over the ranks of a
    torchrun job (points partitioned, NCCL all-reduce of the reduced camera system per lambda trial).
Unit: lambda trials / s (one trial = V*^-1, Schur build, dense solve, updates, FindNewError plus the
amortised projection / Jacobian pass of its LM step).  Timing: device time between CUDA events on
the handle's stream around the whole ptam_bundle_compute call (host LM control included).
Rooflines: the dense solve against the MEASURED f64 peak (profiles/fp64_peak.json: DFMA = DMMA = 37.1 TFLOP/s on
this part), the per-measurement passes against the measured HBM copy bandwidth with SURVEY §8d's bytes.
"""
import json
import time
from pathlib import Path

import numpy as np

from . import synth
from .capi import Bundle

ROOT = Path(__file__).resolve().parent.parent

CONFIGS = {
    "C3": dict(n_cams=50, n_points=5000, n_meas=20000, seed=42),
    "C4": dict(n_cams=500, n_points=100000, n_meas=600000, seed=43),
}


def flops_per_trial(n):
    """dense LDL^T + two triangular solves."""
    return n ** 3 / 3.0 + 2.0 * n ** 2


def fp64_peak():
    """(TFLOP/s, source): the f64 tensor-core (DMMA m8n8k4) burst rate measured on this pool's B200 by
    scripts/fp64_peak.cu (plain DFMA reaches the same 37.1 TFLOP/s: the part has no separate f64 tensor rate)."""
    p = ROOT / "profiles" / "fp64_peak.json"
    try:
        d = json.loads(p.read_text())
        return float(d["dmma_tflops"]), "measured (profiles/fp64_peak.json, scripts/fp64_peak.cu)"
    except Exception:
        return 37.1, "fallback (earlier measurement on this pool)"


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def _run(lib, g, device=0, shard=None, profile=False, **kw):
    b = Bundle(lib, g["width"], g["height"], device=device, **kw) if device is not None else Bundle(lib, g["width"], g["height"], **kw)
    b.add_graph(g)
    if shard is not None:
        b.set_shard(*shard)
    if profile:
        b.set_profiling(True)
    return b


def cpu_leg(cpu_lib, g, trials=2):
    """The reference's (or the oracle's) Bundle::Compute capped at 1 and at `trials` lambda trials: the difference is
    the cost of the extra trials, the rest of the 1-trial run is set-up (its dense [camera][point] LUT and scripts,
    Bundle.cc:558-599) + one trial (time a fixed small number of trials and separate the set-up)."""
    cpu_lib, kind = cpu_lib if isinstance(cpu_lib, tuple) else (cpu_lib, "port")
    t = {}
    done = {}
    for k in (1, trials):
        o = Bundle(cpu_lib, g["width"], g["height"], max_iterations=k)
        o.add_graph(g)
        t0 = time.perf_counter()
        o.Compute()
        t[k] = time.perf_counter() - t0
        done[k] = o.stats().lambda_trials
        o.close()
    per_trial = (t[trials] - t[1]) / max(done[trials] - done[1], 1)
    setup = max(t[1] - per_trial * done[1], 0.0)
    return {"value": 1.0 / per_trial, "unit": "lambda-trials/s", "cores": 1, "kind": kind,
            "setup_ms": setup * 1e3, "ms_per_trial": per_trial * 1e3,
            "value_including_setup_over_20_trials": 20.0 / (setup + 20.0 * per_trial),
            "sample": f"Compute() capped at 1 and at {trials} lambda trials on the same graph ({t[1]:.2f} s, {t[trials]:.2f} s): "
                      f"per-trial cost = the difference, set-up = the rest"}


def bench_ba(prod, device=0, config="C3", reps=3, cpu_lib=None, shard=None, graph=None, cpu_trials=2):
    """shard: None or (rank, world, ncclComm_t from capi.nccl_comm_create).  cpu_lib: an already-loaded CPU library
    exporting the same ABI, or (library, kind) (bench.py passes oracle/_ref or the oracle for the cpu_baseline leg; this package never
    loads it itself)."""
    import torch
    cfg = CONFIGS[config]
    g = graph if graph is not None else synth.make_ba_graph(**cfg)
    n = 6 * int((np.asarray(g["cam_fixed"]) == 0).sum())
    M = len(g["meas_cam"])
    best = None
    runs = []
    for r in range(reps + 1):  # first repetition is the warm-up
        b = _run(prod, g, device, shard, profile=(r == reps))  # last repetition: per-phase events (adds synchronisation, not the timed one)
        b.synchronize()
        ext = torch.cuda.ExternalStream(b.cuda_stream(), device=torch.device("cuda", device))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = b.launch_count()
        t0 = time.perf_counter()
        e0.record(ext)
        acc = b.Compute()
        e1.record(ext)
        b.synchronize()
        wall = time.perf_counter() - t0
        dt = e0.elapsed_time(e1) * 1e-3
        s = b.stats()
        runs.append((acc, s.lambda_trials, s.n_outliers, s.last_error))
        if 0 < r < reps and (best is None or dt < best[0]):
            best = (dt, acc, s.lambda_trials, s.lm_steps, s.n_outliers, b.launch_count() - l0, wall)
        phases = b.phase_times() if r == reps else None
        if r == reps:
            result = (b.GetOutlierMeasurements().copy(), b.get_points(), b.get_cameras())
            prof_wall = wall
        b.close()
    dt, acc, trials, steps, outl, launches, wall = best
    per_call = {k: (v[0] / v[1] if v[1] else None) for k, v in phases.items() if k != "reserved"}
    total = {k: v[0] for k, v in phases.items() if k != "reserved"}
    out = {"workload": f"{config}: Bundle::Compute LM, {cfg['n_cams']} keyframes x {cfg['n_points']} points x {cfg['n_meas']} measurements",
           "value": trials / dt, "unit": "lambda-trials/s", "accepted_steps_per_s": acc / dt, "compute_ms": dt * 1e3,
           "wall_ms": wall * 1e3, "lambda_trials": trials, "accepted": acc, "lm_steps": steps, "outliers": outl,
           "gpu_launches": launches, "reduced_system_n": n,
           "repeat_runs_identical": len(set(runs)) == 1,
           "timing": "CUDA events on the handle's stream around ptam_bundle_compute (host LM control + device phases), best of %d" % (reps - 1),
           "phases_ms_per_call": per_call,
           "phases_calls": {k: int(v[1]) for k, v in phases.items() if k != "reserved"},
           "phases_ms_total_profiled_run": total,
           "phases_sum_over_wall_profiled_run": sum(total.values()) / (prof_wall * 1e3)}
    sol, jac, upd = phases["solve"], phases["jacobian"], phases["update_newerror"]
    peak64, src64 = fp64_peak()
    hbm, srch = hbm_peak()
    rl = {}
    if sol[1]:
        ach = flops_per_trial(n) / (sol[0] / sol[1] * 1e-3) / 1e12
        out["solve_gflops"] = ach * 1e3
        rl["solve"] = {"kernels": "k_ldlt_dag (persistent blocked LDL^T: chain CTA + row-solve / diagonal / DMMA m8n8k4 tile tasks; k_ldlt_step/_update for the tile-bound head of large systems) + k_ldlt_scale/_back", "bound": "tensor_f64",
                       "flops_per_launch_group": flops_per_trial(n), "ms": sol[0] / sol[1], "achieved": ach, "peak": peak64,
                       "unit": "TFLOP/s", "frac": ach / peak64, "peak_source": src64}
    if jac[1]:  # b3 + b5: 320 B per measurement per LM step over the projection + Jacobian phases
        prj = phases["project"]
        ms = jac[0] / jac[1] + (prj[0] / prj[1] if prj[1] else 0.0)
        ach = 320.0 * M / (ms * 1e-3) / 1e9
        rl["project_jacobian"] = {"kernels": "k_ba_project + k_ba_jacobian + k_ba_acc_cam + k_ba_acc_pt", "bound": "hbm",
                                  "alg_bytes": 320.0 * M, "ms": ms, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                                  "peak_source": srch}
    if upd[1]:  # b10: 152 B per measurement per lambda trial
        ach = 152.0 * M / (upd[0] / upd[1] * 1e-3) / 1e9
        rl["new_error"] = {"kernels": "k_ba_cam_update + k_ba_point_update + k_ba_new_error", "bound": "hbm", "alg_bytes": 152.0 * M,
                           "ms": upd[0] / upd[1], "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "peak_source": srch}
    out["roofline"] = rl
    out["_result"] = result
    if cpu_lib is not None:
        out["cpu_baseline"] = cpu_leg(cpu_lib, g, trials=cpu_trials)
    return out
