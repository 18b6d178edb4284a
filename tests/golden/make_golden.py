"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference ships no golden vectors and cannot be built
here, so these fixtures freeze the oracle's outputs on small seeded inputs; the known-answer tests
in tests/test_oracle_known_answers.py are what pins the oracle itself."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from ptam_cg_b200 import synth  # noqa: E402
from ptam_cg_b200.capi import Bundle, Tracker  # noqa: E402
from oracle.binding import detect_with, oracle_lib  # noqa: E402


def tracker_case(lib, frames, poses, kfs, m, start, frame):
    W, H = frames.shape[2], frames.shape[1]
    t = Tracker(lib, W, H, 1)
    for k in kfs:
        t.add_keyframe(k)
    t.set_map(0, m)
    t.set_state(0, pose12=start, msd=0.02)
    r = t.track_frames([frames[frame]])[0]
    p = t.get_points(0)
    out = dict(pose=np.array(r.se3_cam_from_world), attempted=np.array(r.meas_attempted), found=np.array(r.meas_found),
               did_coarse=r.did_coarse, n_sets=np.array([r.n_coarse, r.n_level3, r.n_fine]),
               flags=p["flags"], level=p["level"], v2_found=p["v2_found"], iteration_set=t.get_iteration_set(0),
               templates=t.get_templates(0)[0])
    for l in range(4):
        pix, xy, lut = t.get_level(0, l)
        out[f"corners{l}"] = xy
        out[f"lut{l}"] = lut
    return out


def main():
    lib = oracle_lib()
    W, H = 192, 144
    frames, poses = synth.render_sequence(W, H, 6)
    cam = synth.AtanCamera(W, H)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, lib, W, H), cam, kf_indices=(0, 3), per_level=(120, 60, 30, 15))
    start = synth.perturb_pose(poses[4], np.random.default_rng(11))
    out = tracker_case(lib, frames, poses, kfs, m, start, 4)
    np.savez_compressed(ROOT / "tests/golden/tracker_192x144.npz", frames=frames, poses=poses, start=start,
                        kf_indices=np.array([0, 3]), **{"map_" + k: v for k, v in m.items()}, **out)
    g = synth.make_ba_graph(6, 120, 480, seed=21)
    b = Bundle(lib, g["width"], g["height"])
    b.add_graph(g)
    acc = b.Compute()
    s = b.stats()
    np.savez_compressed(ROOT / "tests/golden/bundle_6x120x480.npz", **{"g_" + k: np.asarray(v) for k, v in g.items()},
                        accepted=acc, lambda_trials=s.lambda_trials, n_outliers=s.n_outliers, sigma_squared=s.sigma_squared,
                        outliers=b.GetOutlierMeasurements(), points=b.get_points(), cameras=b.get_cameras())
    print("golden written")


if __name__ == "__main__":
    main()
