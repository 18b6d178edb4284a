"""Driver of scripts/lab/fast_lab.cu: renders a few synthetic frames (the bench's scene), runs the harness."""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from ptam_cg_b200 import synth

def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 480)
    nf = 8
    t0 = time.time()
    frames, poses = synth.render_sequence(w, h, nf)[:2]
    print("rendered", nf, "frames in %.1f s" % (time.time() - t0), flush=True)
    path = "/tmp/fast_lab_frames.raw"
    np.ascontiguousarray(np.stack(frames).astype(np.uint8)).tofile(path)
    here = os.path.dirname(os.path.abspath(__file__))
    S = int(os.environ.get("LAB_S", "296"))
    sys.exit(subprocess.call([os.path.join(here, "fast_lab"), path, str(w), str(h), str(nf), str(S), "20"]))

if __name__ == "__main__":
    main()
