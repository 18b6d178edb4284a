"""Host-side trace of the pipelined submit/collect loop (pinned host frames): time spent inside each call."""
import os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import torch
import bench
from ptam_cg_b200.capi import Tracker, product_lib

def main():
    S, F, K = 296, 32, 12
    lib = product_lib()
    bench.W, bench.H, bench.FRAME_BYTES = 640, 480, 640 * 480
    def detect_factory():
        det = Tracker(lib, 640, 480, 1)
        def detect(image):
            det.make_keyframes([image])
            return [det.get_level(0, l)[:2] for l in range(4)]
        return detect
    frames, poses, kfs, m = bench.build_workload(detect_factory, F, 20260101)
    trk = Tracker(lib, 640, 480, S)
    for k in kfs:
        trk.add_keyframe(k)
    for s in range(S):
        trk.set_map(s, m)
    offsets = [(5 * s) % (2 * F - 2) for s in range(S)]
    bench.init_streams(trk, poses, offsets, F)
    host_frames = torch.from_numpy(frames).pin_memory()
    base = host_frames.data_ptr()
    tabs = [trk.ptr_array([base + bench.pingpong(o + i, F) * bench.FRAME_BYTES for o in offsets]) for i in range(K + 3)]
    for i in range(3):
        trk.submit_array(tabs[i], 640); trk.collect()
    res = trk.result_buffer()
    log = []
    t0 = time.perf_counter()
    trk.submit_array(tabs[3], 640)
    for j in range(1, K):
        a = time.perf_counter(); trk.submit_array(tabs[3 + j], 640); b = time.perf_counter()
        trk.collect_into(res); c = time.perf_counter()
        log.append((a - t0, b - a, c - b))
    trk.collect_into(res)
    tot = time.perf_counter() - t0
    for a, s_, c_ in log:
        print("t=%7.3f ms  submit %.3f ms  collect %.3f ms" % (a * 1e3, s_ * 1e3, c_ * 1e3))
    print("total %.3f ms for %d steps = %.3f ms/step = %.0f frames/s" % (tot * 1e3, K, tot * 1e3 / K, S * K / tot))

main()
