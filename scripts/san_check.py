"""Small end-to-end pass over every kernel family, meant to be run under compute-sanitizer."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Bundle, Tracker, product_lib
prod = product_lib()
g = synth.make_ba_graph(30, 1500, 7000, seed=3)
b = Bundle(prod, g["width"], g["height"], max_iterations=3)
b.add_graph(g)
print("ba", b.Compute(), b.stats().lambda_trials, len(b.GetOutlierMeasurements()))
W, H = 320, 240
frames, poses = synth.render_sequence(W, H, 6)
cam = synth.AtanCamera(W, H)
det = Tracker(prod, W, H, 1)
def detect(im):
    det.make_keyframes([im])
    return [det.get_level(0, l)[:2] for l in range(4)]
kfs, m = synth.build_map(frames, poses, detect, cam, kf_indices=(0, 3), per_level=(150, 80, 40, 20))
t = Tracker(prod, W, H, 2)
for k in kfs:
    t.add_keyframe(k)
for s in range(2):
    t.set_map(s, m)
    t.set_state(s, pose12=synth.perturb_pose(poses[1], np.random.default_rng(s)), velocity=np.zeros(6), msd=0.02)
for f in (1, 2, 3):
    r = t.track_frames([frames[f], frames[f]])
print("trk", sum(r[0].meas_found), r[0].did_coarse)
print("rest", [len(x[0]) for x in t.keyframe_rest(1)])
kf = t.add_keyframe(frames[0])
t.make_keyframes([frames[0], frames[0]])
cands = [x[1] for x in t.keyframe_rest(0)]
t.make_keyframes([frames[5], frames[5]])
print("epi", [int(t.epipolar_search(0, l, kf, poses[0], 1.0, 0.3, poses[5], 0.1, cands[l])[0].sum()) for l in range(4)])
