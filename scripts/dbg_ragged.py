import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Tracker, product_lib
from oracle.binding import oracle_lib, detect_with
oracle, product = oracle_lib(), product_lib()
frames, poses = synth.render_sequence(640, 480, 48)
cam = synth.AtanCamera(640, 480)
kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle, 640, 480), cam, kf_indices=(0, 12, 24, 36))
big = {k: np.concatenate([v, v]) for k, v in m.items()}
for mp in (3000, 1000):
  for mapx, name in ((big, "big"), (m, "normal")):
    o = Tracker(oracle, 640, 480, 1, use_rotation_estimator=0, max_patches_per_frame=mp)
    p = Tracker(product, 640, 480, 1, use_rotation_estimator=0, max_patches_per_frame=mp)
    start = synth.perturb_pose(poses[6], np.random.default_rng(4))
    for t in (o, p):
        for k in kfs: t.add_keyframe(k)
        t.set_map(0, mapx); t.set_state(0, pose12=start, velocity=np.zeros(6), msd=0.0)
    ro, rp = o.track_frames([frames[6]])[0], p.track_frames([frames[6]])[0]
    po, pp = o.get_points(0), p.get_points(0)
    d = np.flatnonzero(po["flags"] != pp["flags"])
    print(mp, name, "found", sum(ro.meas_found), sum(rp.meas_found), "ndiff", len(d), d[:10], po["flags"][d[:10]], pp["flags"][d[:10]], po["level"][d[:10]])
    print("   sets", ro.n_coarse, ro.n_level3, ro.n_fine, rp.n_coarse, rp.n_level3, rp.n_fine, "pose diff", np.abs(np.array(ro.se3_cam_from_world)-np.array(rp.se3_cam_from_world)).max())
    if len(d):
        i = d[0]
        print("   v2found", po["v2_found"][i], pp["v2_found"][i], "v2image", po["v2_image"][i], pp["v2_image"][i])
