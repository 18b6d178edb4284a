"""Pins the tracker oracle against the reference's OWN code: oracle/_ref/libref_ptam.so holds
src/Tracker.cc, MapMaker.cc, KeyFrame.cc, PatchFinder.cc, ImageProcess.cc, Map.cc, Relocaliser.cc and
ATANCamera.cc of /root/reference, compiled in place (oracle/Makefile.ref) against header stand-ins
for TooN / libCVD / GVars3 (oracle/shim/) and driven through ref_tracker_* (oracle/ref_wrap_tracker.cpp).
Tracker::TrackFrame itself runs here: MakeKeyFrame_Lite, the SmallBlurryImage rotation estimator,
PredictPoseWithMotionModel, TrackMap (PVS, coarse + fine stages, SearchForPoints with the PatchFinder,
CalcPoseUpdate), UpdateMotionModel, AssessTrackingQuality.

With the platform atan on both sides the oracle must follow the reference BIT FOR BIT over a sequence
(pose, velocity, scene depth, per-level counters, per-point flags / levels / found positions /
M-estimator counters, FAST corner lists and row LUTs).  What stays restated is the arithmetic of the
three absent libraries (FAST-10 by its definition, halfSample, transform/sample, convolveGaussian,
SE3/SO3 exp+ln, LDL^T, WLS) and std::random_shuffle = identity (oracle/shim/ref_prelude.h)."""
import numpy as np
import pytest

from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Tracker
from oracle.binding import detect_with, oracle_lib, ref_lib

REF = ref_lib()
pytestmark = pytest.mark.skipif(REF is None or not REF.has("tracker_create"), reason="oracle/_ref not built and /root/reference absent")

IN_IMAGE, IN_PVS, SEARCHED, FOUND, SUBPIX = 1, 2, 4, 8, 16


def _scene(W, H, n_frames, kf_indices, per_level, seed=20260101):
    frames, poses = synth.render_sequence(W, H, n_frames, seed=seed)
    cam = synth.AtanCamera(W, H)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle_lib(), W, H), cam, kf_indices=kf_indices, per_level=per_level)
    return frames, poses, kfs, m


def _tracker(lib, W, H, kfs, m, **params):
    t = Tracker(lib, W, H, 1, **params)
    for k in kfs:
        t.add_keyframe(k)
    t.set_map(0, m)
    return t


def _compare_frame(to, tr, ro, rr, exact=True, tol=0.0):
    """to / tr: oracle / reference trackers after the same TrackFrame; ro / rr their results."""
    eq = (lambda a, b: np.array_equal(a, b)) if exact else (lambda a, b: np.allclose(a, b, rtol=0, atol=tol))
    assert eq(np.array(ro.se3_cam_from_world), np.array(rr.se3_cam_from_world))
    assert list(ro.meas_attempted) == list(rr.meas_attempted) and list(ro.meas_found) == list(rr.meas_found)
    assert list(ro.n_corners) == list(rr.n_corners)
    assert ro.did_coarse == rr.did_coarse
    assert eq(ro.scene_depth_mean, rr.scene_depth_mean) and eq(ro.scene_depth_sigma, rr.scene_depth_sigma)
    so, sr = to.get_state(0), tr.get_state(0)
    assert eq(np.array(so.velocity), np.array(sr.velocity))
    assert eq(so.msd_scaled_velocity_magnitude, sr.msd_scaled_velocity_magnitude)
    assert (so.lost_frames, so.frame, so.just_recovered_so_use_coarse) == (sr.lost_frames, sr.frame, sr.just_recovered_so_use_coarse)
    if not ro.quality_needs_kf_distance:  # that branch consults the map maker (Tracker.cc:1095-1099), left to the caller
        assert so.tracking_quality == sr.tracking_quality
    for l in range(4):
        (_, xo, lo), (_, xr, lr) = to.get_level(0, l), tr.get_level(0, l)
        assert np.array_equal(xo, xr) and np.array_equal(lo, lr)
    po, pr = to.get_points(0), tr.get_points(0)
    pvs = (po["flags"] & IN_PVS) != 0
    # points that did not enter the PVS this frame: off-image, or no usable warp (nSearchLevel == -1)
    assert np.all(((pr["flags"][~pvs] & IN_IMAGE) == 0) | (pr["level"][~pvs] == -1))
    mask = IN_IMAGE | SEARCHED | FOUND | SUBPIX
    assert np.array_equal(po["flags"][pvs] & mask, pr["flags"][pvs] & mask)
    assert np.array_equal(po["level"][pvs], pr["level"][pvs])
    found = pvs & ((po["flags"] & FOUND) != 0)
    assert eq(po["v2_found"][found], pr["v2_found"][found])
    assert eq(po["v2_image"][pvs], pr["v2_image"][pvs])
    assert np.array_equal(po["outliers"], pr["outliers"]) and np.array_equal(po["inliers"], pr["inliers"])
    return int(found.sum())


@pytest.mark.parametrize("use_sbi", [1, 0], ids=["rotation-estimator", "motion-model-only"])
def test_trackframe_sequence_bit_identical_to_reference(use_sbi):
    W, H = 320, 240
    frames, poses, kfs, m = _scene(W, H, 14, (0, 6), (150, 80, 40, 20))
    orc = oracle_lib(libm_atan=True)
    to, tr = (_tracker(lib, W, H, kfs, m, use_rotation_estimator=use_sbi) for lib in (orc, REF))
    start = synth.perturb_pose(poses[2], np.random.default_rng(0))
    for t in (to, tr):
        t.set_state(0, pose12=start, msd=0.02)
    coarse = 0
    for f in range(2, 13):
        ro, rr = to.track_frames([frames[f]])[0], tr.track_frames([frames[f]])[0]
        assert _compare_frame(to, tr, ro, rr) > 100
        coarse += ro.did_coarse
        if use_sbi:
            so, sr = to.get_sbi(0)[0], tr.get_sbi(0)[0]
            assert np.array_equal(so, sr)  # SmallBlurryImage::mimTemplate, float for float
    assert coarse > 0  # both the coarse and the fine-only schedules were exercised


def test_trackframe_640x480_c2_like():
    """BASELINE config C2 scale: 640x480, ~1000-point map."""
    W, H = 640, 480
    frames, poses, kfs, m = _scene(W, H, 6, (0, 4), (600, 250, 100, 50))
    to, tr = (_tracker(lib, W, H, kfs, m) for lib in (oracle_lib(libm_atan=True), REF))
    start = synth.perturb_pose(poses[1], np.random.default_rng(3))
    for t in (to, tr):
        t.set_state(0, pose12=start, msd=0.02)
    for f in (1, 2, 3):
        ro, rr = to.track_frames([frames[f]])[0], tr.track_frames([frames[f]])[0]
        assert _compare_frame(to, tr, ro, rr) > 400


def test_spec_atan_oracle_follows_reference_within_tolerance():
    """The oracle the CUDA product is checked against uses the specified (fdlibm) atan: same decisions,
    poses within 1e-9 of the reference built with the platform atan."""
    W, H = 320, 240
    frames, poses, kfs, m = _scene(W, H, 8, (0, 6), (150, 80, 40, 20))
    to, tr = (_tracker(lib, W, H, kfs, m) for lib in (oracle_lib(), REF))
    start = synth.perturb_pose(poses[2], np.random.default_rng(0))
    for t in (to, tr):
        t.set_state(0, pose12=start, msd=0.02)
    for f in range(2, 7):
        ro, rr = to.track_frames([frames[f]])[0], tr.track_frames([frames[f]])[0]
        assert np.allclose(np.array(ro.se3_cam_from_world), np.array(rr.se3_cam_from_world), rtol=0, atol=1e-9)
        assert list(ro.meas_found) == list(rr.meas_found) and list(ro.meas_attempted) == list(rr.meas_attempted)


def test_lost_tracking_counters_match_reference():
    """Noise frames: nothing is found, tracking quality goes BAD, the lost-frame counter runs, and from
    the third lost frame on the reference attempts recovery instead of tracking (Tracker.cc:133,170-178)."""
    W, H = 320, 240
    frames, poses, kfs, m = _scene(W, H, 8, (0, 6), (150, 80, 40, 20))
    to, tr = (_tracker(lib, W, H, kfs, m) for lib in (oracle_lib(libm_atan=True), REF))
    for t in (to, tr):
        t.set_state(0, pose12=poses[2], msd=0.02)
    rng = np.random.default_rng(5)
    seq = [frames[2], frames[3]] + [rng.integers(0, 256, (H, W), dtype=np.uint8) for _ in range(2)]
    for im in seq:
        ro, rr = to.track_frames([im])[0], tr.track_frames([im])[0]
        _compare_frame(to, tr, ro, rr)
    assert to.get_state(0).lost_frames == 2


def test_refind_in_keyframe_matches_reference():
    """MapMaker::ReFindInSingleKeyFrame / ReFind_Common (MapMaker.cc:943-1040), SURVEY 8f rank 3."""
    W, H = 320, 240
    frames, poses, kfs, m = _scene(W, H, 10, (0, 8), (150, 80, 40, 20))
    to, tr = (_tracker(lib, W, H, kfs, m) for lib in (oracle_lib(libm_atan=True), REF))
    for f in (3, 5):
        for t in (to, tr):
            t.refind_in_keyframes([frames[f]], [poses[f]])
        po, pr = to.get_points(0), tr.get_points(0)
        fo, fr = (po["flags"] & FOUND) != 0, (pr["flags"] & FOUND) != 0
        assert np.array_equal(fo, fr) and fo.sum() > 100
        assert np.array_equal(po["flags"][fo] & SUBPIX, pr["flags"][fo] & SUBPIX)
        assert np.array_equal(po["level"][fo], pr["level"][fo])
        assert np.array_equal(po["v2_found"][fo], pr["v2_found"][fo])


def test_keyframe_rest_matches_reference():
    """KeyFrame::MakeKeyFrame_Rest (KeyFrame.cc:61-82): maximal corners, Shi-Tomasi candidates and scores."""
    W, H = 320, 240
    frames, _ = synth.render_sequence(W, H, 3)
    to, tr = (Tracker(lib, W, H, 1) for lib in (oracle_lib(libm_atan=True), REF))
    n = 0
    rest = []
    for t in (to, tr):
        t.make_keyframes([frames[1]])
        rest.append(t.keyframe_rest(0, 70.0))
    for (mo, co, so), (mr, cr, sr) in zip(*rest):
        assert np.array_equal(mo, mr) and np.array_equal(co, cr) and np.array_equal(so, sr)
        n += len(co)
    assert n > 50


def _epipolar_case(lib, W, H, frames, poses, i_src, i_tgt, depth=(1.0, 0.3), wiggle=0.1):
    """Candidates = Shi-Tomasi candidates of the source frame; returns per level (cand, found, best, sub)."""
    t = Tracker(lib, W, H, 1)
    kf = t.add_keyframe(frames[i_src])
    t.make_keyframes([frames[i_src]])
    rest = t.keyframe_rest(0, 70.0)
    t.make_keyframes([frames[i_tgt]])
    out = []
    for l in range(4):
        cand = rest[l][1]
        out.append((cand,) + t.epipolar_search(0, l, kf, poses[i_src], depth[0], depth[1], poses[i_tgt], wiggle, cand))
    return out


@pytest.mark.parametrize("pair", [(0, 30), (10, 25), (30, 0)], ids=lambda p: f"src{p[0]}-tgt{p[1]}")
def test_epipolar_search_matches_reference(pair):
    """MapMaker::AddPointEpipolar (MapMaker.cc:529-688), SURVEY 8f rank 3: the same candidates accepted,
    the same sub-pixel positions in the target keyframe."""
    W, H = 320, 240  # NB AddPointEpipolar caches UnProject per pixel in a function-local static: one size per process
    frames, poses = synth.render_sequence(W, H, 40)
    o = _epipolar_case(oracle_lib(libm_atan=True), W, H, frames, poses, *pair)
    r = _epipolar_case(REF, W, H, frames, poses, *pair)
    total = 0
    for (co, fo, bo, so), (cr, fr, br, sr) in zip(o, r):
        assert np.array_equal(co, cr) and np.array_equal(fo, fr)
        assert np.array_equal(so, sr)
        total += int(fo.sum())
    assert total > 200


def test_epipolar_search_degenerate_geometry_matches_reference():
    """Same pose for both keyframes (zero baseline: the projected segment collapses) and a depth range behind
    the target: every early exit of AddPointEpipolar."""
    W, H = 320, 240
    frames, poses = synth.render_sequence(W, H, 40)
    for args in [dict(i_src=5, i_tgt=5), dict(i_src=0, i_tgt=30, depth=(-3.0, 0.1), wiggle=-5.0), dict(i_src=0, i_tgt=30, depth=(50.0, 1.0))]:
        o = _epipolar_case(oracle_lib(libm_atan=True), W, H, frames, poses, **args)
        r = _epipolar_case(REF, W, H, frames, poses, **args)
        for (co, fo, bo, so), (cr, fr, br, sr) in zip(o, r):
            assert np.array_equal(fo, fr) and np.array_equal(so, sr)


def test_relocaliser_recovery_matches_reference():
    """Tracker::TrackFrame's recovery branch (Tracker.cc:170-178,196-207; Relocaliser.cc:12-38; SURVEY 8f rank
    4): three noise frames lose tracking, the next real frame is relocalised against the stored keyframes'
    small blurry images and tracked from the recovered pose; then tracking continues normally."""
    W, H = 320, 240
    frames, poses, kfs, m = _scene(W, H, 14, (0, 6), (150, 80, 40, 20))
    to, tr = (_tracker(lib, W, H, kfs, m) for lib in (oracle_lib(libm_atan=True), REF))
    for t in (to, tr):
        for i, k in enumerate((0, 6)):
            t.set_keyframe_pose(i, poses[k])
        t.set_state(0, pose12=poses[2], msd=0.02)
    rng = np.random.default_rng(5)
    noise = [rng.integers(0, 256, (H, W), dtype=np.uint8) for _ in range(3)]
    seq = [frames[2], frames[3]] + noise + [frames[7], frames[8], frames[9]]
    modes = []
    for im in seq:
        ro, rr = to.track_frames([im])[0], tr.track_frames([im])[0]
        assert ro.recovery == rr.recovery
        if ro.recovery:
            assert ro.reloc_keyframe == rr.reloc_keyframe
        _compare_frame(to, tr, ro, rr)
        modes.append(ro.recovery)
    assert modes == [0, 0, 0, 0, 0, 1, 0, 0]
    assert to.get_state(0).lost_frames == 0 and sum(ro.meas_found) > 100  # back on track


def test_relocaliser_failure_leaves_the_tracker_alone():
    """A frame nothing like any keyframe: the ESM score stays above Reloc2.MaxScore and nothing is done."""
    W, H = 320, 240
    frames, poses, kfs, m = _scene(W, H, 8, (0, 6), (150, 80, 40, 20))
    to, tr = (_tracker(lib, W, H, kfs, m) for lib in (oracle_lib(libm_atan=True), REF))
    for t in (to, tr):
        for i, k in enumerate((0, 6)):
            t.set_keyframe_pose(i, poses[k])
        t.set_state(0, pose12=poses[2], msd=0.02)
    rng = np.random.default_rng(9)
    checker = ((np.indices((H, W)).sum(0) // 8) % 2 * 255).astype(np.uint8)
    seq = [frames[2]] + [rng.integers(0, 256, (H, W), dtype=np.uint8) for _ in range(3)] + [checker, 255 - checker]
    modes = []
    for im in seq:
        ro, rr = to.track_frames([im])[0], tr.track_frames([im])[0]
        assert ro.recovery == rr.recovery
        so, sr = to.get_state(0), tr.get_state(0)
        assert np.array_equal(np.array(so.se3_cam_from_world), np.array(sr.se3_cam_from_world))
        assert (so.lost_frames, so.frame) == (sr.lost_frames, sr.frame)
        modes.append(ro.recovery)
    assert modes[:4] == [0, 0, 0, 0] and modes[4] in (1, 2)


@pytest.mark.parametrize("seed", [20260102, 20260103, 20260104])
def test_other_scenes_bit_identical_to_reference(seed):
    """The C5 stream seeds (seed 20260101 + g): other textures, other corner sets, other candidate
    lists — the oracle must still follow the reference's TrackFrame bit for bit, with other settings too."""
    W, H = 320, 240
    frames, poses, kfs, m = _scene(W, H, 10, (0, 5), (120, 60, 30, 15), seed=seed)
    prm = dict(use_rotation_estimator=seed & 1, mestimator=seed % 3, coarse_min_velocity=0.0 if seed % 2 else 0.006)
    to, tr = (_tracker(lib, W, H, kfs, m, **prm) for lib in (oracle_lib(libm_atan=True), REF))
    start = synth.perturb_pose(poses[1], np.random.default_rng(seed))
    for t in (to, tr):
        t.set_state(0, pose12=start, msd=0.02)
    for f in range(1, 8):
        ro, rr = to.track_frames([frames[f]])[0], tr.track_frames([frames[f]])[0]
        assert _compare_frame(to, tr, ro, rr) > 60
