"""GPU parity: CUDA path T (through the C-ABI) against the CPU oracle on the same inputs.
Bar: bit-exact pyramid pixels, FAST corner lists (order included), row LUTs, search levels,
template bytes, found flags and coarse patch positions; poses / sub-pixel positions to 1e-9."""
import numpy as np
import pytest

from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Tracker, PT_FOUND, PT_SUBPIX

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-9      # absolute, SE3 entries (rotation entries ~1, translations ~1)
SUBPIX_TOL = 1e-6    # pixels (float32 arithmetic inside IterateSubPix, order of summation differs)


def _compare_levels(a: Tracker, b: Tracker, stream=0):
    for l in range(4):
        pa, ca, la = a.get_level(stream, l)
        pb, cb, lb = b.get_level(stream, l)
        assert np.array_equal(pa, pb), f"level {l} pixels differ"
        assert ca.shape == cb.shape and np.array_equal(ca, cb), f"level {l} corners differ ({len(ca)} vs {len(cb)})"
        assert np.array_equal(la, lb), f"level {l} LUT differs"


@pytest.mark.parametrize("shape", [(640, 480), (1280, 720), (321, 243), (64, 64), (200, 67)])
def test_keyframe_lite_matches_oracle(oracle, product, shape):
    w, h = shape
    rng = np.random.default_rng(5)
    tex = synth.make_texture(seed=3)
    imgs = [
        tex[100:100 + h, 200:200 + w].copy(),
        rng.integers(0, 256, (h, w), dtype=np.uint8),
        np.full((h, w), 77, np.uint8),
        (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8),
    ]
    o, p = Tracker(oracle, w, h), Tracker(product, w, h)
    for im in imgs:
        o.make_keyframes([im]); p.make_keyframes([im])
        _compare_levels(o, p)


def test_keyframe_batch_streams(oracle, product, seq640):
    frames, _ = seq640
    S = 5
    o, p = Tracker(oracle, 640, 480, S), Tracker(product, 640, 480, S)
    ims = [frames[3 * i] for i in range(S)]
    o.make_keyframes(ims); p.make_keyframes(ims)
    for s in range(S):
        _compare_levels(o, p, s)


def _setup(lib, kfs, m, S=1, **prm):
    """The bit-exact per-frame comparisons below run with the velocity-only motion model: with the
    SmallBlurryImage rotation estimator on, the predicted pose carries the ~1e-13 difference of the
    estimator's parallel sums, and the bilinear->byte truncation of the coarse templates can then flip
    a byte (the sensitivity test_libm_atan_variant_* documents).  The estimator has its own parity test
    (test_sbi_rotation_estimator_matches_oracle) and is on by default everywhere else."""
    prm.setdefault("use_rotation_estimator", 0)
    t = Tracker(lib, 640, 480, S, **prm)
    for k in kfs:
        t.add_keyframe(k)
    for s in range(S):
        t.set_map(s, m)
    return t


def _compare_frame(o, p, ro, rp, stream=0):
    for f in ("meas_attempted", "meas_found", "n_corners", "n_pvs"):
        assert list(getattr(ro, f)) == list(getattr(rp, f)), f
    for f in ("did_coarse", "n_coarse", "n_level3", "n_fine", "tracking_quality", "quality_needs_kf_distance"):
        assert getattr(ro, f) == getattr(rp, f), f
    assert np.array_equal(o.get_iteration_set(stream), p.get_iteration_set(stream))
    po, pp = o.get_points(stream), p.get_points(stream)
    assert np.array_equal(po["level"], pp["level"])
    assert np.array_equal(po["flags"], pp["flags"])
    to, so = o.get_templates(stream)
    tp, sp = p.get_templates(stream)
    assert np.array_equal(to, tp), "template bytes differ"
    assert np.array_equal(so, sp)
    found = (po["flags"] & PT_FOUND) != 0
    sub = (po["flags"] & PT_SUBPIX) != 0
    coarse_only = found & ~sub
    assert np.array_equal(po["v2_found"][coarse_only], pp["v2_found"][coarse_only]), "coarse patch positions differ"
    np.testing.assert_allclose(pp["v2_found"][sub], po["v2_found"][sub], atol=SUBPIX_TOL, rtol=0)
    np.testing.assert_allclose(pp["v2_image"], po["v2_image"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(np.array(rp.se3_cam_from_world), np.array(ro.se3_cam_from_world), atol=POSE_TOL, rtol=0)
    np.testing.assert_allclose(rp.scene_depth_mean, ro.scene_depth_mean, atol=1e-9)
    np.testing.assert_allclose(rp.scene_depth_sigma, ro.scene_depth_sigma, atol=1e-7)
    assert np.array_equal(po["outliers"], pp["outliers"])
    assert np.array_equal(po["inliers"], pp["inliers"])


def test_track_frame_from_identical_state(oracle, product, seq640, map640):
    """Every frame: both sides start from the oracle's state (per-step parity from identical state)."""
    frames, poses = seq640
    kfs, m = map640
    o, p = _setup(oracle, kfs, m), _setup(product, kfs, m)
    rng = np.random.default_rng(7)
    start = synth.perturb_pose(poses[5], rng)
    o.set_state(0, pose12=start); p.set_state(0, pose12=start)
    n_coarse_frames = 0
    for f in range(5, 30):
        st = o.get_state(0)
        p.set_state(0, state=st)
        ro = o.track_frames([frames[f]])[0]
        rp = p.track_frames([frames[f]])[0]
        _compare_frame(o, p, ro, rp)
        n_coarse_frames += ro.did_coarse
        so, sp = o.get_state(0), p.get_state(0)
        np.testing.assert_allclose(np.array(sp.velocity), np.array(so.velocity), atol=1e-9)
        np.testing.assert_allclose(sp.msd_scaled_velocity_magnitude, so.msd_scaled_velocity_magnitude, atol=1e-9)
        assert sum(ro.meas_found) > 200
    assert n_coarse_frames > 5  # the coarse stage was exercised


def test_default_settings_sequence_template_bytes(oracle, product, seq640, map640):
    """The same per-frame comparison with the SHIPPED settings (Tracker.UseRotationEstimator = 1, what the bench and
    the reference's default run use): every frame starts from the oracle's state, and all integer outputs are
    asserted — levels, flags, iteration sets, per-level counts and the 8x8 template bytes + sums.  The predicted
    pose then carries the rotation estimator's ~1e-13 summation-order difference (k_sbi sums its 15 ESM
    accumulators by shuffle reduction, the reference pixel by pixel), and the bilinear -> byte truncation of
    CVD::transform turns that into a different byte wherever a sample lies within ~1e-11 of an integer: over these
    25 frames x 1000 points x 64 bytes a handful of bytes differ, each by exactly 1, and stay in the per-point
    template cache until the warp moves.  Everything else (levels, flags, found sets, counts) stays equal."""
    frames, poses = seq640
    kfs, m = map640
    o, p = _setup(oracle, kfs, m, use_rotation_estimator=1), _setup(product, kfs, m, use_rotation_estimator=1)
    start = synth.perturb_pose(poses[5], np.random.default_rng(7))
    o.set_state(0, pose12=start); p.set_state(0, pose12=start)
    n_tmpl = 0
    worst_diff = 0
    for f in range(5, 30):
        p.set_state(0, state=o.get_state(0))
        ro = o.track_frames([frames[f]])[0]
        rp = p.track_frames([frames[f]])[0]
        for fld in ("meas_attempted", "meas_found", "n_corners", "n_pvs"):
            assert list(getattr(ro, fld)) == list(getattr(rp, fld)), (f, fld)
        assert (ro.did_coarse, ro.n_coarse, ro.n_level3, ro.n_fine) == (rp.did_coarse, rp.n_coarse, rp.n_level3, rp.n_fine)
        assert np.array_equal(o.get_iteration_set(0), p.get_iteration_set(0))
        po, pp = o.get_points(0), p.get_points(0)
        assert np.array_equal(po["level"], pp["level"]) and np.array_equal(po["flags"], pp["flags"])
        to, so = o.get_templates(0)
        tp, sp = p.get_templates(0)
        diff = to.astype(np.int32) - tp.astype(np.int32)
        n_diff = int((diff != 0).sum())
        worst_diff = max(worst_diff, n_diff)
        assert n_diff <= 8 and (n_diff == 0 or np.abs(diff).max() == 1), (f, n_diff, np.abs(diff).max())
        assert np.abs(so - sp).max() <= 2 * 255 * 8
        n_tmpl += int((so[:, 1] > 0).sum())
        np.testing.assert_allclose(np.array(rp.se3_cam_from_world), np.array(ro.se3_cam_from_world), atol=POSE_TOL, rtol=0)
    assert n_tmpl > 5000
    print("default settings: most template bytes differing in one frame:", worst_diff, "of", 64 * len(to))


def test_track_sequence_free_running(oracle, product, seq640, map640):
    """Whole run without re-synchronising.  A 1e-15 pose difference (f64 summation order) can flip
    a discrete decision (template refresh at 0.07, ir() truncation, sub-pixel convergence) a few
    frames later, after which the two runs use slightly different measurement sets; both must keep
    tracking the same trajectory, so the poses are compared loosely (well inside tracker noise,
    which is ~5e-4 against ground truth on this sequence)."""
    frames, poses = seq640
    kfs, m = map640
    o, p = _setup(oracle, kfs, m), _setup(product, kfs, m)
    start = synth.perturb_pose(poses[5], np.random.default_rng(7))
    o.set_state(0, pose12=start); p.set_state(0, pose12=start)
    worst = 0.0
    for f in range(5, 40):
        ro = o.track_frames([frames[f]])[0]
        rp = p.track_frames([frames[f]])[0]
        worst = max(worst, np.abs(np.array(rp.se3_cam_from_world) - np.array(ro.se3_cam_from_world)).max())
        assert abs(sum(rp.meas_found) - sum(ro.meas_found)) <= 10
    assert worst < 2e-4, worst
    Rt, tt = synth.se3_from12(poses[39])
    assert np.abs(np.array(rp.se3_cam_from_world)[9:] - tt).max() < 5e-3
    assert np.abs(np.array(ro.se3_cam_from_world)[9:] - tt).max() < 5e-3


def test_batched_streams_match_single(oracle, product, seq640, map640):
    frames, poses = seq640
    kfs, m = map640
    S = 4
    o, p = _setup(oracle, kfs, m, S), _setup(product, kfs, m, S)
    rng = np.random.default_rng(11)
    for s in range(S):
        st = synth.perturb_pose(poses[6 + 2 * s], rng)
        o.set_state(s, pose12=st, msd=0.01 * s); p.set_state(s, pose12=st, msd=0.01 * s)
    ims = [frames[6 + 2 * s] for s in range(S)]
    ro, rp = o.track_frames(ims), p.track_frames(ims)
    for s in range(S):
        _compare_frame(o, p, ro[s], rp[s], s)


def test_settings_variants(oracle, product, seq640, map640):
    frames, poses = seq640
    kfs, m = map640
    for prm in (dict(disable_coarse=1), dict(max_patches_per_frame=150), dict(coarse_max=10, coarse_min=5),
                dict(mestimator=1), dict(mestimator=2), dict(use_constant_velocity=0)):
        o, p = _setup(oracle, kfs, m, **prm), _setup(product, kfs, m, **prm)
        start = synth.perturb_pose(poses[8], np.random.default_rng(3))
        for t in (o, p):
            t.set_state(0, pose12=start, msd=0.02, just_recovered=(1 if "coarse_max" in prm else 0))
        ro, rp = o.track_frames([frames[8]])[0], p.track_frames([frames[8]])[0]
        _compare_frame(o, p, ro, rp)


def test_device_resident_frames(product, oracle, seq640, map640):
    import torch
    frames, poses = seq640
    kfs, m = map640
    S = 3
    o, p = _setup(oracle, kfs, m, S), _setup(product, kfs, m, S)
    start = synth.perturb_pose(poses[10], np.random.default_rng(1))
    for s in range(S):
        o.set_state(s, pose12=start); p.set_state(s, pose12=start)
    ims = [frames[10], frames[10], frames[11]]
    d = torch.from_numpy(np.stack(ims)).cuda()
    torch.cuda.synchronize()
    rp = p.track_frames_device(d.data_ptr(), 640 * 480, 640, want_results=True)
    ro = o.track_frames(ims)
    for s in range(S):
        _compare_frame(o, p, ro[s], rp[s], s)


def test_empty_map_and_no_cuda_fallback(product):
    t = Tracker(product, 640, 480, 2)
    im = np.zeros((480, 640), np.uint8)
    r = t.track_frames([im, im])
    assert sum(r[0].meas_attempted) == 0 and r[0].tracking_quality == 0
    assert t.launch_count() > 0


@pytest.mark.gpu
def test_pipelined_submit_collect_matches_blocking(product, seq640, map640):
    """ptam_tracker_submit_frames / _collect (two batches in flight) gives the same results as the
    blocking ptam_tracker_track_frames on the same frames."""
    frames, poses = seq640
    kfs, m = map640
    S = 3
    trackers = []
    for _ in range(2):
        t = Tracker(product, 640, 480, S)
        for k in kfs:
            t.add_keyframe(k)
        for s in range(S):
            t.set_map(s, m)
            t.set_state(s, pose12=synth.perturb_pose(poses[2 + s], np.random.default_rng(s)), velocity=np.zeros(6), msd=0.0)
        trackers.append(t)
    a, b = trackers
    batches = [[np.ascontiguousarray(frames[2 + s + i]) for s in range(S)] for i in range(5)]
    ref = [a.track_frames(bt) for bt in batches]
    out = []
    b.submit_ptrs([im.ctypes.data for im in batches[0]], 640)
    for i in range(1, 5):
        b.submit_ptrs([im.ctypes.data for im in batches[i]], 640)
        out.append(b.collect())
    out.append(b.collect())
    with pytest.raises(Exception):
        b.collect()
    for r_step, o_step in zip(ref, out):
        for r, o in zip(r_step, o_step):
            assert list(r.se3_cam_from_world) == list(o.se3_cam_from_world)
            assert list(r.meas_found) == list(o.meas_found) and list(r.n_corners) == list(o.n_corners)


@pytest.mark.gpu
def test_pipelined_submit_device_matches_blocking(product, seq640, map640):
    """ptam_tracker_submit_frames_device / _collect (device-resident frames, image work of batch i+1 beside the
    fine pose iterations of batch i) gives bit-identical results to the blocking call, n_corners included."""
    import torch
    frames, poses = seq640
    kfs, m = map640
    S = 5
    trackers = []
    for _ in range(2):
        t = Tracker(product, 640, 480, S)
        for k in kfs:
            t.add_keyframe(k)
        for s in range(S):
            t.set_map(s, m)
            t.set_state(s, pose12=synth.perturb_pose(poses[2 + s], np.random.default_rng(s)), velocity=np.zeros(6), msd=0.0)
        trackers.append(t)
    a, b = trackers
    n = 8
    batches = [[np.ascontiguousarray(frames[2 + s + i]) for s in range(S)] for i in range(n)]
    ref = [a.track_frames(bt) for bt in batches]
    dev = [torch.from_numpy(np.stack(bt)).cuda() for bt in batches]
    torch.cuda.synchronize()
    out = []
    b.submit_device(dev[0].data_ptr(), 640 * 480, 640)
    for i in range(1, n):
        b.submit_device(dev[i].data_ptr(), 640 * 480, 640)
        out.append(b.collect())
    out.append(b.collect())
    # a blocking call after the pipelined ones sees their state
    last = [np.ascontiguousarray(frames[2 + s + n]) for s in range(S)]
    ra, rb = a.track_frames(last), b.track_frames(last)
    for r_step, o_step in zip(ref + [ra], out + [rb]):
        for r, o in zip(r_step, o_step):
            assert list(r.se3_cam_from_world) == list(o.se3_cam_from_world)
            assert list(r.meas_found) == list(o.meas_found) and list(r.n_corners) == list(o.n_corners)
            assert list(r.meas_attempted) == list(o.meas_attempted)


@pytest.mark.gpu
def test_track_frames_1280x720(oracle, product):
    """BASELINE config C5 geometry: full TrackFrame parity at 1280x720 (map built at that size)."""
    from oracle.binding import detect_with
    Wd, Hd = 1280, 720
    frames, poses = synth.render_sequence(Wd, Hd, 6)
    cam = synth.AtanCamera(Wd, Hd)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle, Wd, Hd), cam, kf_indices=(0, 3), per_level=(300, 150, 60, 30))
    trk = []
    for lib in (oracle, product):
        t = Tracker(lib, Wd, Hd, 1, use_rotation_estimator=0)
        for k in kfs:
            t.add_keyframe(k)
        t.set_map(0, m)
        t.set_state(0, pose12=synth.perturb_pose(poses[1], np.random.default_rng(1)), velocity=np.zeros(6), msd=0.0)
        trk.append(t)
    o, p = trk
    for f in (1, 2, 4):
        ro, rp = o.track_frames([frames[f]])[0], p.track_frames([frames[f]])[0]
        _compare_levels(o, p)
        po, pp = o.get_points(0), p.get_points(0)
        assert np.array_equal(po["flags"], pp["flags"]) and np.array_equal(po["level"], pp["level"])
        assert list(ro.meas_found) == list(rp.meas_found) and sum(rp.meas_found) > 100
        assert np.allclose(np.array(ro.se3_cam_from_world), np.array(rp.se3_cam_from_world), atol=1e-9)
        p.set_state(0, state=o.get_state(0))  # continue from identical state


@pytest.mark.gpu
def test_sbi_rotation_estimator_matches_oracle(oracle, product, seq640, map640):
    """k_sbi: the small blurry image is bit-identical to the oracle's, the ESM rotation estimate and
    the pose predicted from it agree to 1e-9 (parallel sums), frame after frame."""
    frames, poses = seq640
    kfs, m = map640
    trk = []
    for lib in (oracle, product):
        t = Tracker(lib, 640, 480, 1)
        for k in kfs:
            t.add_keyframe(k)
        t.set_map(0, m)
        t.set_state(0, pose12=synth.perturb_pose(poses[3], np.random.default_rng(9)), velocity=np.zeros(6), msd=0.0)
        trk.append(t)
    o, p = trk
    for f in range(3, 9):
        ro, rp = o.track_frames([frames[f]])[0], p.track_frames([frames[f]])[0]
        (to, roto, so), (tp, rotp, sp) = o.get_sbi(0), p.get_sbi(0)
        assert np.array_equal(to, tp), "small blurry image differs"
        assert np.allclose(roto, rotp, atol=1e-9) and abs(so - sp) <= 1e-9 * max(1.0, abs(so))
        assert np.allclose(np.array(ro.se3_cam_from_world), np.array(rp.se3_cam_from_world), atol=1e-9)
        p.set_state(0, state=o.get_state(0))
    # the estimator can be switched off (the velocity-only model of the kernel-only numbers)
    t0 = Tracker(product, 640, 480, 1, use_rotation_estimator=0)
    assert t0.params.use_rotation_estimator == 0


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(640, 480), (200, 67)])
def test_keyframe_rest_matches_oracle(oracle, product, shape):
    """KeyFrame::MakeKeyFrame_Rest: vMaxCorners and Shi-Tomasi candidates bit-exact (positions, order,
    f64 scores) on textured, random and flat images."""
    w, h = shape
    rng = np.random.default_rng(11)
    tex = synth.make_texture(seed=3)
    imgs = [tex[100:100 + h, 200:200 + w].copy(), rng.integers(0, 256, (h, w), dtype=np.uint8), np.full((h, w), 9, np.uint8)]
    o, p = Tracker(oracle, w, h), Tracker(product, w, h)
    for im in imgs:
        o.make_keyframes([im]); p.make_keyframes([im])
        for thr in (70.0, 400.0):
            ro, rp = o.keyframe_rest(0, thr), p.keyframe_rest(0, thr)
            for l in range(4):
                assert np.array_equal(ro[l][0], rp[l][0]), f"vMaxCorners differ on level {l}"
                assert np.array_equal(ro[l][1], rp[l][1]) and np.array_equal(ro[l][2], rp[l][2]), f"candidates differ on level {l}"
    assert len(rp[0][0]) == 0  # flat image: nothing


@pytest.mark.gpu
def test_ragged_streams_and_large_found_set(oracle, product, seq640, map640):
    """Streams of one batch with different map sizes (empty, small, and a 2000-point map whose found set
    exceeds the 1024-point shared-memory working set of k_pose, so the global-memory path runs), with
    MaxPatchesPerFrame raised accordingly."""
    frames, poses = seq640
    kfs, m = map640
    n = len(m["src_kf"])
    empty = {k: v[:0] for k, v in m.items()}
    small = {k: v[:300] for k, v in m.items()}
    big = {k: np.concatenate([v, v]) for k, v in m.items()}  # every point twice: 2n points
    maps = [empty, small, big]
    o = Tracker(oracle, 640, 480, 3, use_rotation_estimator=0, max_patches_per_frame=3000)
    p = Tracker(product, 640, 480, 3, use_rotation_estimator=0, max_patches_per_frame=3000)
    start = synth.perturb_pose(poses[6], np.random.default_rng(4))
    for t in (o, p):
        for k in kfs:
            t.add_keyframe(k)
        for s in range(3):
            t.set_map(s, maps[s])
            t.set_state(s, pose12=start, velocity=np.zeros(6), msd=0.0)
    ims = [frames[6]] * 3
    ro, rp = o.track_frames(ims), p.track_frames(ims)
    assert sum(rp[0].meas_attempted) == 0
    assert sum(rp[2].meas_found) > 1024, sum(rp[2].meas_found)
    for s in range(3):
        _compare_frame(o, p, ro[s], rp[s], s)


@pytest.mark.gpu
def test_scale_change_and_lost_frames(oracle, product, seq640, map640):
    """Search-level selection / template rejection under a scale change (camera 2.2x higher and 0.55x
    lower than the keyframes: det of the warp leaves [0.25, 3] for many points, PatchFinder.cc:70-81),
    then frames of noise: tracking quality goes BAD and mnLostFrames counts up (Tracker.cc:1062-1107)."""
    frames, poses = seq640
    kfs, m = map640
    tex = synth.make_texture()
    cam = synth.AtanCamera(640, 480)
    for scale in (2.2, 0.55):
        R, t = synth.se3_from12(poses[8])
        c = -R.T @ t
        c2 = c.copy(); c2[2] *= scale
        pose = synth.se3_to12(R, -R @ c2)
        im = synth.render_frame(tex, cam, pose)
        o, p = _setup(oracle, kfs, m), _setup(product, kfs, m)
        for trk in (o, p):
            trk.set_state(0, pose12=synth.perturb_pose(pose, np.random.default_rng(2)), velocity=np.zeros(6), msd=0.0)
        ro, rp = o.track_frames([im])[0], p.track_frames([im])[0]
        _compare_frame(o, p, ro, rp)
        lv = o.get_points(0)["level"]
        assert len(set(lv[lv >= 0].tolist())) >= 2      # several search levels in play
        assert ((o.get_points(0)["flags"] & 32) != 0).sum() > 0   # some templates rejected
    rng = np.random.default_rng(0)
    o, p = _setup(oracle, kfs, m), _setup(product, kfs, m)
    start = synth.perturb_pose(poses[5], np.random.default_rng(7))
    for trk in (o, p):
        trk.set_state(0, pose12=start, velocity=np.zeros(6), msd=0.0)
    for k in range(4):
        noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)
        ro, rp = o.track_frames([noise])[0], p.track_frames([noise])[0]
        _compare_frame(o, p, ro, rp)
        so, sp = o.get_state(0), p.get_state(0)
        assert (so.tracking_quality, so.lost_frames) == (sp.tracking_quality, sp.lost_frames)
        p.set_state(0, state=so)
    assert so.tracking_quality == 0 and so.lost_frames >= 3


@pytest.mark.gpu
def test_refind_in_keyframes_matches_oracle(oracle, product, seq640, map640):
    """MapMaker::ReFindInSingleKeyFrame: two keyframes at once (one per stream), all
    map points searched with radius 4 around their projection; found / sub-pixel flags, levels and
    coarse positions bit-exact, sub-pixel positions to 1e-9; most visible points are re-found close to
    where they project."""
    frames, poses = seq640
    kfs, m = map640
    S = 2
    o, p = _setup(oracle, kfs, m, S), _setup(product, kfs, m, S)
    ims = [frames[7], frames[30]]
    kposes = np.stack([poses[7], poses[30]])
    o.refind_in_keyframes(ims, kposes); p.refind_in_keyframes(ims, kposes)
    for s in range(S):
        _compare_levels(o, p, s)
        po, pp = o.get_points(s), p.get_points(s)
        mask = PT_FOUND | PT_SUBPIX | 2
        assert np.array_equal(po["flags"] & mask, pp["flags"] & mask)
        assert np.array_equal(po["level"], pp["level"])
        found = (po["flags"] & PT_FOUND) != 0
        sub = (po["flags"] & PT_SUBPIX) != 0
        assert np.array_equal(po["v2_found"][found & ~sub], pp["v2_found"][found & ~sub])
        np.testing.assert_allclose(pp["v2_found"][found & sub], po["v2_found"][found & sub], atol=1e-9, rtol=0)
        in_pvs = (po["flags"] & 2) != 0
        assert found.sum() > 0.5 * in_pvs.sum() > 100
        err = np.abs(po["v2_found"][found] - po["v2_image"][found]).max(axis=1)
        assert (err <= 4.0 * 2.0 ** po["level"][found] + 2.0 ** po["level"][found]).all()
        assert (sub == (found & (po["level"] > 0))).all()   # sub-pixel exactly on levels > 0
    # the tracking state of the handle is untouched
    assert np.array_equal(np.array(p.get_state(0).se3_cam_from_world), np.array(Tracker(product, 640, 480, 1).get_state(0).se3_cam_from_world))
