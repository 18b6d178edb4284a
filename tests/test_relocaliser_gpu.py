"""Recovery branch of Tracker::TrackFrame on the device (k_kf_sbi, k_reloc; Relocaliser.cc:12-38,
Tracker.cc:170-178,196-207) against the CPU oracle (itself pinned bit-exact against the
reference's Tracker.cc / Relocaliser.cc in tests/test_ref_pin_tracker.py) and against the reference itself."""
import numpy as np
import pytest

from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Tracker, product_lib
from oracle.binding import detect_with, oracle_lib, ref_lib

pytestmark = pytest.mark.gpu
FOUND, SUBPIX, PVS = 8, 16, 2


def _setup(libs, W, H, n_frames, kf_idx, per_level, _unused=1):
    frames, poses = synth.render_sequence(W, H, n_frames)
    cam = synth.AtanCamera(W, H)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle_lib(), W, H), cam, kf_indices=kf_idx, per_level=per_level)
    trk = []
    for lib, S in libs:
        t = Tracker(lib, W, H, S)
        for i, k in enumerate(kfs):
            t.add_keyframe(k)
        for i, k in enumerate(kf_idx):
            t.set_keyframe_pose(i, poses[k])
        for s in range(S):
            t.set_map(s, m)
        trk.append(t)
    return frames, poses, trk


def _same(tp, to, rp, ro, sp=0, so=0, pose_tol=1e-9):
    assert rp.recovery == ro.recovery and rp.reloc_keyframe == ro.reloc_keyframe
    assert list(rp.meas_attempted) == list(ro.meas_attempted) and list(rp.meas_found) == list(ro.meas_found)
    assert rp.did_coarse == ro.did_coarse
    np.testing.assert_allclose(np.array(rp.se3_cam_from_world), np.array(ro.se3_cam_from_world), rtol=0, atol=pose_tol)
    stp, sto = tp.get_state(sp), to.get_state(so)
    assert (stp.lost_frames, stp.frame, stp.tracking_quality, stp.just_recovered_so_use_coarse) == \
           (sto.lost_frames, sto.frame, sto.tracking_quality, sto.just_recovered_so_use_coarse)
    np.testing.assert_allclose(np.array(stp.velocity), np.array(sto.velocity), rtol=0, atol=pose_tol)
    pp, po = tp.get_points(sp), to.get_points(so)
    pvs = (po["flags"] & PVS) != 0
    assert np.array_equal(pp["flags"][pvs], po["flags"][pvs]) and np.array_equal(pp["level"][pvs], po["level"][pvs])


def test_lost_then_relocalised_matches_oracle():
    W, H = 320, 240
    frames, poses, (tp, to) = _setup([(product_lib(), 1), (oracle_lib(), 1)], W, H, 14, (0, 6), (150, 80, 40, 20))
    for t in (tp, to):
        t.set_state(0, pose12=poses[2], msd=0.02)
    rng = np.random.default_rng(5)
    seq = [frames[2], frames[3]] + [rng.integers(0, 256, (H, W), dtype=np.uint8) for _ in range(3)] + [frames[7], frames[8], frames[9]]
    modes = []
    for im in seq:
        rp, ro = tp.track_frames([im])[0], to.track_frames([im])[0]
        _same(tp, to, rp, ro)
        if ro.recovery:
            np.testing.assert_allclose(rp.reloc_score, ro.reloc_score, rtol=1e-9)
        modes.append(rp.recovery)
        tp.set_state(0, state=to.get_state(0))  # identical start for the next frame
    assert modes == [0, 0, 0, 0, 0, 1, 0, 0] and sum(rp.meas_found) > 100


def test_mixed_batch_only_the_lost_stream_recovers():
    """Three streams in one batch: one tracking normally, one lost (relocalised), one lost on a frame no keyframe
    resembles... each must equal a single-stream oracle run of its own history."""
    W, H = 320, 240
    prod, orc = product_lib(), oracle_lib()
    frames, poses, (tp, o0, o1, o2) = _setup([(prod, 3), (orc, 1), (orc, 1), (orc, 1)], W, H, 14, (0, 6), (150, 80, 40, 20), 1)
    # _setup built the product with S = 3 only for the first entry
    rng = np.random.default_rng(11)
    noise = [rng.integers(0, 256, (H, W), dtype=np.uint8) for _ in range(3)]
    checker = ((np.indices((H, W)).sum(0) // 8) % 2 * 255).astype(np.uint8)
    hist = [
        [frames[2], frames[3], frames[4], frames[5], frames[6]],
        [frames[2]] + noise + [frames[7]],
        [frames[2]] + noise + [checker],
    ]
    for s, o in enumerate((o0, o1, o2)):
        tp.set_state(s, pose12=poses[2], msd=0.02)
        o.set_state(0, pose12=poses[2], msd=0.02)
    for f in range(5):
        rp = tp.track_frames([hist[s][f] for s in range(3)])
        for s, o in enumerate((o0, o1, o2)):
            ro = o.track_frames([hist[s][f]])[0]
            _same(tp, o, rp[s], ro, sp=s, so=0)
            tp.set_state(s, state=o.get_state(0))
    assert [r.recovery for r in rp][:2] == [0, 1]


def test_product_recovery_follows_the_reference():
    ref = ref_lib()
    if ref is None or not ref.has("tracker_set_keyframe_pose"):
        pytest.skip("oracle/_ref/libref_ptam.so not present")
    W, H = 320, 240
    frames, poses, (tp, tr) = _setup([(product_lib(), 1), (ref, 1)], W, H, 14, (0, 6), (150, 80, 40, 20))
    for t in (tp, tr):
        t.set_state(0, pose12=poses[2], msd=0.02)
    rng = np.random.default_rng(5)
    seq = [frames[2]] + [rng.integers(0, 256, (H, W), dtype=np.uint8) for _ in range(3)] + [frames[7], frames[8]]
    for im in seq:
        rp, rr = tp.track_frames([im])[0], tr.track_frames([im])[0]
        assert rp.recovery == rr.recovery
        assert list(rp.meas_found) == list(rr.meas_found)
        np.testing.assert_allclose(np.array(rp.se3_cam_from_world), np.array(rr.se3_cam_from_world), rtol=0, atol=5e-9)
        tp.set_state(0, state=tr.get_state(0))
    assert sum(rp.meas_found) > 100
