"""clock64 phase breakdown of k_pose (debug build scripts/lab/libptam_b200_dbg.so, -DPTAM_POSE_CLOCKS): the bench
workload (640x480, ~1000-point map, S streams), a few frames, cycles of CTA 0 per phase."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import bench
from ptam_cg_b200 import capi
from ptam_cg_b200.capi import Tracker, Lib

def main():
    S = int(os.environ.get("LAB_S", "296"))
    lib = Lib(os.path.join(ROOT, "scripts", "lab", "libptam_b200_dbg.so"), "ptam_")
    bench.W, bench.H, bench.FRAME_BYTES = 640, 480, 640 * 480
    F = 32

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
    fn = lib.cdll.ptam_debug_pose_clocks
    fn.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    out = (C.c_longlong * 32)()
    n = 6
    for i in range(3 + n):
        if i == 3:
            trk.synchronize(); fn(None, 1)
        trk.track_frames([frames[bench.pingpong(o + i, F)] for o in offsets])
    trk.synchronize()
    fn(out, 0)
    names = {0: "per-point (non-linear its)", 1: "per-point (linear its)", 2: "sigma select", 3: "WLS accumulate", 4: "transpose-sum + smem",
             5: "cross-warp sum + 6x6 solve", 6: "se3 exp + pose", 8: "compaction + gather", 9: "(iterations total marker)", 10: "write-back + motion model/quality"}
    for stage, off in (("fine", 0), ("coarse", 16)):
        tot = sum(out[off + k] for k in range(11) if k != 9)
        print("cycles of CTA 0 per frame, %s k_pose, S = %d" % (stage, S))
        for k in sorted(names):
            if k == 9: continue
            print("  %-34s %9.0f  %5.1f %%" % (names[k], out[off + k] / n, 100.0 * out[off + k] / max(tot, 1)))
        print("  total %.0f cycles = %.1f us at 1.965 GHz" % (tot / n, tot / n / 1965.0))

if __name__ == "__main__":
    main()
