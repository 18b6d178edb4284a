"""One epipolar search per level and one relocalisation frame, for ncu captures of k_epi_* / k_reloc."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Tracker, product_lib
W, H = 640, 480
frames, poses = synth.render_sequence(W, H, 40)
prod = product_lib()
t = Tracker(prod, W, H, 1)
kf = t.add_keyframe(frames[0])
t.set_keyframe_pose(kf, poses[0])
t.make_keyframes([frames[0]])
cands = [r[1] for r in t.keyframe_rest(0, 70.0)]
t.make_keyframes([frames[30]])
for l in range(4):
    f, b, s = t.epipolar_search(0, l, kf, poses[0], 1.0, 0.3, poses[30], 0.1, cands[l])
    print("level", l, "candidates", len(cands[l]), "found", int(f.sum()))
st = t.get_state(0)
st.lost_frames = 3
t.set_state(0, state=st)
r = t.track_frames([frames[2]])[0]
print("recovery", r.recovery, "keyframe", r.reloc_keyframe, "score", r.reloc_score)
