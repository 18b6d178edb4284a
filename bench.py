#!/usr/bin/env python
"""bench.py — tracker frames/s (+ BA lambda-trials/s) on B200, with roofline and CPU baseline.

Workload at N=1 (BASELINE.json configs[1]): 640x480, 4-level pyramid, ~1000-point map, full
TrackFrame (MakeKeyFrame_Lite + motion model + TrackMap + quality) for a batch of S independent
streams per step.  One "step" = one TrackFrame for every stream of the batch.

  value : whole-job frames/s with the frames already resident in HBM (device-resident batches,
          CUDA events on the library's stream, max over ranks).
  e2e   : the same through the C-ABI call a user makes, frames in pinned HOST memory, H2D of the
          frames and D2H of the per-stream results inside the timed region.
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md §Measurement.

N>1 (torchrun): every rank runs its own S streams on its own GPU (replicas, no collective on the
data path — SURVEY.md §8e path T), weak scaling; rank 0 prints the one JSON line.
`--impl reference` times the reference's own Tracker::TrackFrame (oracle/_ref: its sources compiled
against TooN/libCVD/GVars3 header stand-ins; the oracle port when that library is absent) on all
host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H = 640, 480
FRAME_BYTES = W * H


def workload_name():
    if (W, H) == (640, 480):
        return "C2: 640x480 4-level pyramid, ~1000-point map, full TrackFrame"
    if (W, H) == (1280, 720):
        return "C5: 1280x720 4-level pyramid, ~1000-point map, full TrackFrame, independent streams per GPU"
    return f"{W}x{H} 4-level pyramid, ~1000-point map, full TrackFrame"


def pingpong(i, n):
    p = i % (2 * n - 2)
    return p if p < n else 2 * n - 2 - p


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d.get("hbm_gbs", 6650.0)), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """Polls NVML during the timed region (the region is tens of ms: nvidia-smi -lms is too slow)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self._active = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            if not self._active:
                time.sleep(0.0005)
                continue
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if os.environ.get("PTAM_BENCH_NO_SAMPLER"):
            self.nv = None
        """Start the polling thread ahead of the timed region (thread start-up and the first NVML call are
        not free); it only samples between __enter__ and __exit__."""
        if self.nv and self._t is None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __enter__(self):
        self.start()
        self._active = True
        return self

    def __exit__(self, *a):
        self._active = False
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_workload(detect_factory, n_frames, seed):
    from ptam_cg_b200 import synth
    frames, poses = synth.render_sequence(W, H, n_frames, seed=seed)
    cam = synth.AtanCamera(W, H)
    q = n_frames // 4
    kfs, m = synth.build_map(frames, poses, detect_factory(), cam, kf_indices=(0, q, 2 * q, 3 * q))
    return frames, poses, kfs, m


def init_streams(trk, poses, offsets, n_frames, rng_seed=7):
    from ptam_cg_b200 import synth
    rng = np.random.default_rng(rng_seed)
    for s, off in enumerate(offsets):
        trk.set_state(s, pose12=synth.perturb_pose(poses[pingpong(off, n_frames)], rng),
                      velocity=np.zeros(6), msd=0.0, depth_mean=1.0)


def pin_to_gpu_numa_node(gpu_index):
    """One process per GPU: run on the cores NVML calls local to this GPU, so that the pinned frame buffers (first
    touch) and the copy submissions sit on the GPU's own NUMA node / PCIe root.  Returns what was done, for the line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return {"pinned": False, "why": "no NVML-local core is allowed to this process"}
        os.sched_setaffinity(0, allowed)
        return {"pinned": True, "cores": f"{allowed[0]}-{allowed[-1]} ({len(allowed)})"}
    except Exception as e:  # NVML absent, cpuset restrictions ...
        return {"pinned": False, "why": repr(e)[:120]}


def cpu_reference_lib():
    """(library, kind) for the CPU legs: oracle/_ref/libref_ptam.so — the reference's own Tracker.cc /
    Bundle.cc ... compiled against header stand-ins (oracle/Makefile.ref; prebuilt, it travels with the
    snapshot) — when present (kind "reference"), else the oracle port (kind "port")."""
    from oracle.binding import oracle_lib
    so = ROOT / "oracle" / "_ref" / "libref_ptam.so"
    if so.exists():
        try:
            from ptam_cg_b200.capi import Lib
            lib = Lib(so, "ref_")
            if lib.has("tracker_create") and lib.has("bundle_create"):
                return lib, "reference"
        except Exception:
            pass
    return oracle_lib(), "port"


def run_cpu_baseline(cpu, kfs, m, frames, poses, budget_s=10.0):
    """The reference's Tracker::TrackFrame (or the oracle port), one core (the reference runs the
    tracker on one thread): single stream, consecutive frames, for ~budget_s seconds."""
    from ptam_cg_b200.capi import Tracker
    orc, kind = cpu
    t = Tracker(orc, W, H, 1)
    for k in kfs:
        t.add_keyframe(k)
    t.set_map(0, m)
    init_streams(t, poses, [0], len(frames))
    for i in range(3):
        t.track_frames([frames[pingpong(i, len(frames))]])
    if kind == "port" and hasattr(orc.cdll, "orc_tracker_cpu_split"):
        import ctypes
        orc.cdll.orc_tracker_cpu_split((ctypes.c_double * 6)(), 1)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        t.track_frames([frames[pingpong(3 + n, len(frames))]])
        n += 1
    dt = time.perf_counter() - t0
    out = {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": kind,
           "sample": f"{n} consecutive TrackFrame calls, 1 stream, same frames/map as the GPU arm, {dt:.1f}s"}
    if kind == "port" and hasattr(orc.cdll, "orc_tracker_cpu_split"):
        # BASELINE config C1: where one CPU frame goes (the oracle port carries timers; the reference's own code does not)
        import ctypes
        sp = (ctypes.c_double * 6)()
        orc.cdll.orc_tracker_cpu_split(sp, 1)
        fr = max(sp[5], 1.0)
        names = ("pyramid_ms", "fast10_lut_ms", "sbi_rotation_ms", "search_for_points_ms", "calc_pose_update_ms")
        split = {k: 1e3 * sp[i] / fr for i, k in enumerate(names)}
        split["other_ms"] = 1e3 * dt / max(n, 1) - sum(split.values())
        out["c1_cost_split_per_frame"] = split
    if kind == "reference" and budget_s > 2.0:
        # the oracle port (faster than the reference built on stand-in libraries) beside it, for scale
        from oracle.binding import oracle_lib
        p = run_cpu_baseline((oracle_lib(), "port"), kfs, m, frames, poses, budget_s=budget_s / 2)
        out["oracle_port"] = {"value": p["value"], "unit": "frames/s", "cores": 1, "sample": p["sample"]}
        if "c1_cost_split_per_frame" in p:  # BASELINE config C1: pyramid / FAST / search / GN of one CPU frame
            out["oracle_port"]["c1_cost_split_per_frame"] = p["c1_cost_split_per_frame"]
    return out


def reference_arm(args, rank, world):
    """--impl reference: the reference's own Tracker::TrackFrame compiled here (oracle/_ref), or the CPU
    oracle port when that library is absent, on all host threads (rank 0 only)."""
    if rank != 0:
        return
    from oracle.binding import oracle_lib, detect_with
    from ptam_cg_b200.capi import Tracker
    orc, kind = cpu_reference_lib()
    n_frames = args.frames
    frames, poses, kfs, m = build_workload(lambda: detect_with(Tracker, oracle_lib(), W, H), n_frames, 20260101)
    threads = os.cpu_count() or 1
    per_step = 4  # frames per thread per step (bounded sample of the S-stream batch)
    trackers = []
    for th in range(threads):
        t = Tracker(orc, W, H, 1)
        for k in kfs:
            t.add_keyframe(k)
        t.set_map(0, m)
        init_streams(t, poses, [(3 * th) % (n_frames - 1)], n_frames, rng_seed=7 + th)
        trackers.append(t)
    pos = [0] * threads

    def work(th):
        t = trackers[th]
        off = (3 * th) % (n_frames - 1)
        for _ in range(per_step):
            t.track_frames([frames[pingpong(off + pos[th], n_frames)]])
            pos[th] += 1

    def step():
        ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        for x in ths:
            x.start()
        for x in ths:
            x.join()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    n = args.steps * threads * per_step
    val = n / dt
    sample = f"{threads} threads x {per_step} TrackFrame per step (independent streams), {n} frames in {dt:.1f}s"
    print(json.dumps({
        "impl": "reference", "metric": "tracker frames/sec", "value": val, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32/f64",
        "data": "synthetic",
        "config": {"workload": workload_name(), "map_points": int(len(m["src_kf"]))},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), file=JSON_OUT, flush=True)


def _claim_stdout():
    """The reference's own code (oracle/_ref) writes progress text to fd 1 ("Waiting for mapmaker to die..").
    Keep the real stdout for the one JSON line and send everything else written to fd 1 to stderr."""
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=296, help="independent tracker streams per GPU (batch)")
    ap.add_argument("--frames", type=int, default=64, help="distinct synthetic frames per trajectory")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--res", default="640x480", help="frame size; 1280x720 = BASELINE config C5 (one independent stream set per GPU)")
    ap.add_argument("--no-ba", action="store_true")
    ap.add_argument("--ncu-step", action="store_true",
                    help="one extra TrackFrame batch between cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--no-ba-large", action="store_true", help="skip the C4 bundle adjustment (500 x 100k x 600k)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global W, H, FRAME_BYTES
    W, H = (int(v) for v in args.res.lower().split("x"))
    FRAME_BYTES = W * H

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    numa = pin_to_gpu_numa_node(local) if world > 1 else None
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from ptam_cg_b200.capi import Tracker, product_lib
    from ptam_cg_b200 import capi
    prod = product_lib()
    S, K, Wm, F = args.streams, args.steps, args.warmup, args.frames

    def detect_factory():
        det_trk = Tracker(prod, W, H, 1, device=local)

        def detect(image):
            det_trk.make_keyframes([image])
            return [det_trk.get_level(0, l)[:2] for l in range(4)]
        return detect

    # weak scaling = the same work on every GPU: the C2 replicas all track the seed-20260101 sequence (their
    # streams, maps and states are their own); C5 (--res 1280x720) uses seed 20260101 + g for GPU g,
    # whose scenes differ in corner count by ~15 %, which the max-over-ranks time then reflects
    c5 = (W, H) == (1280, 720)
    frames, poses, kfs, m = build_workload(detect_factory, F, 20260101 + (rank if c5 else 0))
    trk = Tracker(prod, W, H, S, device=local)
    for k in kfs:
        trk.add_keyframe(k)
    for s in range(S):
        trk.set_map(s, m)
    offsets = [(5 * s) % (2 * F - 2) for s in range(S)]
    n_steps = Wm + K

    # ---- device-resident batches: step i reads its own S x 307 KB batch (never touched before) ----
    frames_dev = torch.from_numpy(frames).cuda()
    need = n_steps * S * FRAME_BYTES
    if need > 48e9:
        raise SystemExit("steps*streams too large for the resident frame set")
    batches = []
    for i in range(n_steps):
        idx = torch.tensor([pingpong(o + i, F) for o in offsets], device="cuda")
        batches.append(frames_dev.index_select(0, idx).contiguous())
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(trk.cuda_stream(), device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ================= value: frames resident in HBM =================
    # ptam_tracker_submit_frames_device / _collect: the public pipelined call for device-resident frames (two
    # batches in flight; the image kernels of batch i+1 run beside the last pose iterations of batch i).  The
    # collect (a 66 KB result read-back per batch) is inside the timed region; the device time is taken with
    # events on the handle's stream, whose last item of a batch is that read-back.
    init_streams(trk, poses, offsets, F)
    for i in range(Wm):
        trk.submit_device(batches[i].data_ptr(), FRAME_BYTES, W)
        trk.collect(want_results=False)
    trk.synchronize()
    sampler = ClockSampler(local).start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = trk.launch_count()
    with sampler as clk:
        ev0.record(ext)
        th0 = time.perf_counter()
        trk.submit_device(batches[Wm].data_ptr(), FRAME_BYTES, W)
        th_submit = 0.0
        for i in range(Wm + 1, n_steps):
            ts = time.perf_counter()
            trk.submit_device(batches[i].data_ptr(), FRAME_BYTES, W)
            th_submit += time.perf_counter() - ts
            trk.collect(want_results=False)
        trk.collect(want_results=False)
        host_launch_ms = th_submit * 1e3 / max(K - 1, 1)  # host time to queue one step (the device runs behind)
        ev1.record(ext)
        trk.synchronize()
        torch.cuda.synchronize()
    launches = trk.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    barrier()
    last = trk.track_frames_device(batches[n_steps - 1].data_ptr(), FRAME_BYTES, W, want_results=True)
    found = float(np.mean([sum(r.meas_found) for r in last]))
    attempted = float(np.mean([sum(r.meas_attempted) for r in last]))
    n_corners = float(np.mean([sum(r.n_corners) for r in last]))
    n_cand = float(np.mean([r.n_candidates for r in last]))
    n_searched = float(np.mean([r.n_coarse + r.n_level3 + r.n_fine for r in last]))
    tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
    per_rank = None
    if world > 1:
        # every rank's own device time and median SM clock, so that a slow replica is visible in the line
        mine = torch.tensor([ms, float(clk.summary()["sm_mhz"] or 0.0)], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [float(a[0].item()) / K for a in allr], "sm_mhz": [float(a[1].item()) for a in allr]}
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * S * K / (ms_max * 1e-3)

    # ================= e2e: pinned host frames through the C-ABI, results back every step ==========
    host_frames = torch.from_numpy(frames).pin_memory()
    base = host_frames.data_ptr()
    init_streams(trk, poses, offsets, F)

    def frame_ptrs(i):
        return [base + pingpong(o + i, F) * FRAME_BYTES for o in offsets]

    for i in range(Wm):  # warm-up through the same pipelined calls (first use allocates the landing buffers)
        trk.submit_ptrs(frame_ptrs(i), W)
        trk.collect()
    barrier()
    # pipelined public API: submit (H2D of this step's frames + kernels) / collect (D2H of its results);
    # at most two steps in flight, so the copy of step i+1 overlaps the kernels of step i.  The host
    # pointer tables are built beforehand (a capture loop would own them); results land in two buffers.
    ptr_tabs = [trk.ptr_array(frame_ptrs(i)) for i in range(Wm, n_steps)]
    res_bufs = [trk.result_buffer(), trk.result_buffer()]
    t0 = time.perf_counter()
    trk.submit_array(ptr_tabs[0], W)
    for j in range(1, K):
        trk.submit_array(ptr_tabs[j], W)
        res = trk.collect_into(res_bufs[j & 1])
    res = trk.collect_into(res_bufs[K & 1])
    trk.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    e2e_per_rank = None
    if world > 1:
        alle = [torch.zeros_like(te) for _ in range(world)]
        dist.all_gather(alle, te)
        e2e_per_rank = [1e3 * float(a.item()) / K for a in alle]
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    # the bus under the e2e number: pinned host -> device copy rate of this rank, all ranks copying at once
    probe_n = min(host_frames.numel(), 256 << 20)
    probe_src = host_frames.view(-1)[:probe_n]
    probe_dst = torch.empty(probe_n, dtype=torch.uint8, device="cuda")
    barrier()
    h2d_best = 0.0
    for _ in range(3):
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _r in range(max(1, (256 << 20) // probe_n)):
            probe_dst.copy_(probe_src, non_blocking=True)
        pe1.record()
        torch.cuda.synchronize()
        h2d_best = max(h2d_best, max(1, (256 << 20) // probe_n) * probe_n / (pe0.elapsed_time(pe1) * 1e-3) / 1e9)
    del probe_dst
    e2e_value = world * S * K / float(te.item())
    import ctypes
    d2h = S * ctypes.sizeof(capi.TrackResult)

    # ================= per-kernel device times (separate pass, events around every launch) =========
    init_streams(trk, poses, offsets, F)
    for i in range(Wm):
        trk.track_frames_device(batches[i].data_ptr(), FRAME_BYTES, W)
    trk.set_profiling(True)
    for i in range(Wm, n_steps):
        trk.track_frames_device(batches[i].data_ptr(), FRAME_BYTES, W)
    kt = trk.kernel_times()
    trk.set_profiling(False)
    if args.ncu_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        trk.track_frames_device(batches[Wm].data_ptr(), FRAME_BYTES, W)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    peak, peak_src = measured_peaks()
    pyr_px = sum((W >> l) * (H >> l) for l in range(4))
    alg = {  # algorithmic bytes per launch
        # pyramid + FAST of level 0 in one kernel: reads level 0, writes levels 1..3 and the level-0 corner mask
        "k_fast2_l0": S * (pyr_px + W * H // 8),
        # FAST of levels 1..3: reads them back (L2 hits in practice), writes their masks
        "k_fast2_l123": S * ((pyr_px - W * H) + (pyr_px - W * H) // 8),
        "k_compact": S * (pyr_px // 8 + 8 * n_corners + 4 * sum(H >> l for l in range(4))),
        # 64 B template per searched point + (64 B window + 8 B corner) per candidate;
        # coarse and fine launches share the frame's candidate count pro rata to their point counts
        "k_search_fine": S * (64 * attempted + 72 * n_cand),
        # a10: 136 B per found point per Gauss-Newton iteration, ten iterations
        "k_pose_fine": S * 10 * 136 * found,
    }
    per_kernel = {}
    for k, (tot, n) in kt.items():
        if n:
            avg = tot / n
            e = {"avg_ms": avg, "launches": int(n)}
            if k in alg:
                e["alg_bytes"] = float(alg[k])
                e["gbs"] = alg[k] / (avg * 1e-3) / 1e9
                e["frac_hbm"] = e["gbs"] / peak
            per_kernel[k] = e
    step_kernel_ms = sum(v["avg_ms"] for v in per_kernel.values())
    top = max(per_kernel, key=lambda k: per_kernel[k]["avg_ms"])
    traffic = None
    tj = ROOT / "profiles" / "traffic.json"
    if tj.exists() and (W, H) == (640, 480):  # DRAM bytes per launch from the committed ncu --set full capture, scaled to this batch size
        try:
            t = json.loads(tj.read_text())
            if top in t.get("kernels", {}):
                traffic = t["kernels"][top] * S / float(t["streams"])
        except Exception:
            traffic = None
    rl = {"kernel": top, "bound": "hbm", "achieved": per_kernel[top].get("gbs"), "peak": peak, "unit": "GB/s",
          "frac": per_kernel[top].get("frac_hbm"), "traffic": traffic, "peak_source": peak_src,
          "share_of_step": per_kernel[top]["avg_ms"] / step_kernel_ms}
    a1_ms = sum(per_kernel[k]["avg_ms"] for k in ("k_fast2_l0", "k_fast2_l123", "k_compact") if k in per_kernel)
    a1_bytes = S * (pyr_px + 4 * sum(H >> l for l in range(4)) + 8 * n_corners)  # 411 600 + 8 N_c at 640x480
    rl_a1 = {"kernels": "k_fast2_l0+k_fast2_l123+k_compact (SURVEY a1: pyramid+FAST+LUT)", "alg_bytes": a1_bytes,
             "achieved": a1_bytes / (a1_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
             "frac": a1_bytes / (a1_ms * 1e-3) / 1e9 / peak}

    # ================= single-stream latency (S=1, host frames, for the record) ====================
    t1 = Tracker(prod, W, H, 1, device=local)
    for k in kfs:
        t1.add_keyframe(k)
    t1.set_map(0, m)
    init_streams(t1, poses, [0], F)
    for i in range(5):
        t1.track_frames_ptrs([base + pingpong(i, F) * FRAME_BYTES], W)
    t0 = time.perf_counter()
    nlat = 50
    for i in range(5, 5 + nlat):
        t1.track_frames_ptrs([base + pingpong(i, F) * FRAME_BYTES], W)
    lat_ms = (time.perf_counter() - t0) / nlat * 1e3

    out = {
        "metric": "tracker frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
        "warmup": Wm, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32/f64", "data": "synthetic",
        "config": {"workload": workload_name(),
                   "streams_per_gpu": S, "map_points": int(len(m["src_kf"])), "frames_per_step": world * S,
                   "replicas": ("seed 20260101 + g on GPU g" if c5 else "every GPU tracks its own streams of the seed-20260101 sequence (per-GPU work fixed)"),
                   "l2": f"inputs > L2: every step reads a distinct {S}x{FRAME_BYTES} B batch out of a "
                         f"{n_steps * S * FRAME_BYTES / 1e6:.0f} MB resident set",
                   "mean_found_per_frame": found, "mean_attempted_per_frame": attempted,
                   "mean_corners_per_frame": n_corners, "mean_zmssd_candidates_per_frame": n_cand},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": S * FRAME_BYTES,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * float(te.item()) / K,
                "api": "ptam_tracker_submit_frames / ptam_tracker_collect, pinned host frames, 2 steps in flight",
                "h2d_gbs_achieved": world * S * FRAME_BYTES * K / float(te.item()) / 1e9 / world,
                "h2d_gbs_ceiling_this_rank": h2d_best,
                "note": "per-GPU H2D rate of the e2e run beside a plain pinned-memory copy measured in the same process (all ranks copying at once)"},
        "gpu_launches": int(launches), "host_launch_ms_per_step": host_launch_ms,
        "clocks": clk.summary(),
        "roofline": rl, "roofline_a1_group": rl_a1, "kernels": per_kernel,
        "single_stream_latency_ms": lat_ms,
    }
    if per_rank:
        per_rank["e2e_ms_per_step"] = e2e_per_rank
        out["per_rank"] = per_rank
        out["config"]["cpu_affinity_rank0"] = numa

    # ================= BA (configs C3 / C4) =================
    # N = 1: Bundle::Compute on C3 and C4 on this GPU.  N > 1: C4 sharded over all ranks (points
    # partitioned, NCCL reduction of the reduced camera system per lambda trial) - collective, so
    # every rank runs it; rank 0 also runs the same graph on its own GPU alone and reports the parity
    # of the sharded result against it.
    if not args.no_ba and prod.has("bundle_create"):
        ba = {}
        try:
            from ptam_cg_b200.bench_ba import bench_ba
            cpu_lib = None
            if rank == 0 and not args.no_cpu_baseline:
                cpu_lib = cpu_reference_lib()
            if world == 1:
                ba["C3"] = bench_ba(prod, local, "C3", reps=4, cpu_lib=cpu_lib)
                if not args.no_ba_large:
                    ba["C4"] = bench_ba(prod, local, "C4", reps=3, cpu_lib=cpu_lib, cpu_trials=2)
                for v in ba.values():
                    v.pop("_result", None)
            else:
                from ptam_cg_b200 import synth
                from ptam_cg_b200.bench_ba import CONFIGS
                g4 = synth.make_ba_graph(**CONFIGS["C4"])
                uid = torch.tensor(list(capi.nccl_unique_id(prod) if rank == 0 else bytes(capi.NCCL_UNIQUE_ID_BYTES)),
                                   dtype=torch.uint8, device="cuda")
                dist.broadcast(uid, 0)
                comm = capi.nccl_comm_create(prod, local, rank, world, bytes(uid.cpu().tolist()))
                r = bench_ba(prod, local, "C4", reps=3, shard=(rank, world, comm), graph=g4)
                prod.fn("nccl_comm_destroy")(comm)
                mine = torch.tensor([r["compute_ms"]], device="cuda", dtype=torch.float64)
                allc = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(allc, mine)
                r["per_rank_compute_ms"] = [float(a.item()) for a in allc]
                r["sharding"] = (f"points partitioned over {world} ranks; per lambda trial one reduction of the packed lower triangle of S "
                                 f"({r['reduced_system_n']}^2 / 2 f64) + vE over NCCL")
                res_sh = r.pop("_result")
                if rank == 0:  # the same graph on this GPU alone: the sharded run must reproduce it
                    one = bench_ba(prod, local, "C4", reps=2, graph=g4)
                    res_1 = one.pop("_result")
                    r["parity_vs_single_gpu"] = {
                        "trials_equal": one["lambda_trials"] == r["lambda_trials"], "accepted_equal": one["accepted"] == r["accepted"],
                        "outliers_equal": bool(np.array_equal(res_1[0], res_sh[0])),
                        "max_pt": float(np.abs(res_1[1] - res_sh[1]).max()), "max_cam": float(np.abs(res_1[2] - res_sh[2]).max()),
                        "single_gpu_lambda_trials_per_s": one["value"], "single_gpu_compute_ms": one["compute_ms"]}
                    ph, ph1 = r["phases_ms_per_call"], one["phases_ms_per_call"]
                    gain = {k: (ph1[k] or 0.0) - (ph[k] or 0.0) for k in ph if k in ph1}
                    r["limiter"] = ("replicated dense solve %.2f ms of a %.2f ms lambda trial; sharding saves %.2f ms per trial in the "
                                    "per-measurement phases and pays %.2f ms in the reduction and %.2f ms in the distributed select"
                                    % (ph["solve"] or 0.0, r["compute_ms"] / max(r["lambda_trials"], 1),
                                       sum(gain[k] for k in ("jacobian", "schur", "update_newerror", "project", "vinv_init") if k in gain),
                                       ph["allreduce"] or 0.0, (ph["select"] or 0.0) - (ph1["select"] or 0.0)))
                ba["C4_sharded"] = r
        except Exception as e:  # the tracker line must still be printed
            import traceback
            ba["error"] = repr(e) + " | " + traceback.format_exc()[-600:]
        if rank == 0:
            out["ba"] = ba

    if rank == 0 and not args.no_cpu_baseline:
        out["cpu_baseline"] = run_cpu_baseline(cpu_reference_lib(), kfs, m, frames, poses)
    if rank == 0:
        print(json.dumps(out), file=JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
