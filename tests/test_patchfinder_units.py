"""Unit entry points ptam_patch_search_batch / ptam_patch_get_results / ptam_pose_update (reference
include/PatchFinder.h:54-98, src/Tracker.cc:867-1005).

CPU (-m "not gpu"): the oracle's twins against the reference's OWN PatchFinder / Tracker::SearchForPoints /
Tracker::CalcPoseUpdate (oracle/_ref hooks ref_patch_*, ref_pose_update) — bit for bit with the platform atan.
GPU (-m gpu): the CUDA entries against the oracle — integer results (levels, template bytes and sums, found flags,
coarse positions) bit-exact, warp matrices bit-equal, sub-pixel positions to 1e-6 px, pose updates to 1e-9."""
import numpy as np
import pytest

from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Tracker
from oracle.binding import oracle_lib, ref_lib, detect_with

W, H = 320, 240


@pytest.fixture(scope="module")
def scene():
    frames, poses = synth.render_sequence(W, H, 10)
    cam = synth.AtanCamera(W, H)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle_lib(), W, H), cam, kf_indices=(0, 4), per_level=(200, 100, 50, 25))
    return frames, poses, kfs, m


def _tracker(lib, kfs, m, S=1):
    t = Tracker(lib, W, H, S)
    for k in kfs:
        t.add_keyframe(k)
    for s in range(S):
        t.set_map(s, m)
    return t


def _unit_run(t, frame, pose, rng_range, subpix):
    t.make_keyframes([frame] * t.S)
    t.patch_search_batch(np.tile(np.asarray(pose).reshape(1, 12), (t.S, 1)), rng_range, subpix)
    r = t.patch_results(0)
    tm, sums = t.get_templates(0)
    return r, tm, sums


CASES = [(10, 0), (30, 8), (5, 3)]


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built and /root/reference absent")
@pytest.mark.parametrize("case", CASES)
def test_oracle_units_bit_identical_to_reference_classes(scene, case):
    frames, poses, kfs, m = scene
    rng_range, subpix = case
    o, r = _tracker(oracle_lib(libm_atan=True), kfs, m), _tracker(ref_lib(), kfs, m)
    pose = synth.perturb_pose(poses[6], np.random.default_rng(3))
    for frame_idx in (6, 7):   # the second call reuses cached templates where the warp moved < 0.07 (PatchFinder.cc:103-110)
        ro, to, so = _unit_run(o, frames[frame_idx], pose, rng_range, subpix)
        rr, tr, sr = _unit_run(r, frames[frame_idx], pose, rng_range, subpix)
        for k in ("level", "template_bad", "found", "subpix"):
            assert np.array_equal(ro[k], rr[k]), k
        assert np.array_equal(ro["warp_inverse"], rr["warp_inverse"])
        assert np.array_equal(ro["pos"], rr["pos"])
        assert np.array_equal(to, tr) and np.array_equal(so, sr)
        assert ro["found"].sum() > 20
        for override, mark in ((0.0, False), (16.0, True)):
            mo, no = o.pose_update(override, mark)
            mr, nr = r.pose_update(override, mark)
            assert np.array_equal(no, nr) and np.array_equal(mo, mr)
        assert np.array_equal(o.get_points(0)["outliers"], r.get_points(0)["outliers"])


def test_oracle_pose_update_recovers_a_known_motion(scene):
    """Noise-free check of the unit pose update: measurements found at pose T, projected at exp(-xi) T: one
    CalcPoseUpdate moves most of the way back (the 100 I prior and Tukey weights keep it from being exact)."""
    frames, poses, kfs, m = scene
    o = _tracker(oracle_lib(), kfs, m)
    xi = np.array([0.004, -0.003, 0.002, 0.001, -0.002, 0.0015])
    start = synth.se3_to12(*synth.se3_mul(synth.se3_exp(xi), synth.se3_from12(poses[6])))
    o.make_keyframes([frames[6]])
    o.patch_search_batch(np.asarray(start).reshape(1, 12), 12, 8)
    mu, nf = o.pose_update()
    assert nf[0] > 60
    after = synth.se3_to12(*synth.se3_mul(synth.se3_exp(mu[0]), synth.se3_from12(start)))
    assert np.abs(np.asarray(after) - np.asarray(poses[6])).max() < 0.5 * np.abs(np.asarray(start) - np.asarray(poses[6])).max()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_units_match_oracle(oracle, product, scene, case):
    frames, poses, kfs, m = scene
    rng_range, subpix = case
    S = 3
    o, p = _tracker(oracle, kfs, m, S), _tracker(product, kfs, m, S)
    rng = np.random.default_rng(4)
    for frame_idx in (6, 7, 7):
        ps = np.array([synth.perturb_pose(poses[frame_idx], rng) for _ in range(S)])
        fr = [frames[frame_idx]] * S
        o.make_keyframes(fr); p.make_keyframes(fr)
        o.patch_search_batch(ps, rng_range, subpix); p.patch_search_batch(ps, rng_range, subpix)
        for s in range(S):
            ro, rp = o.patch_results(s), p.patch_results(s)
            for k in ("level", "template_bad", "found", "subpix"):
                assert np.array_equal(ro[k], rp[k]), (k, s)
            assert np.array_equal(ro["warp_inverse"], rp["warp_inverse"])          # mm2WarpInverse, bit for bit
            to, so = o.get_templates(s)
            tp, sp = p.get_templates(s)
            assert np.array_equal(to, tp) and np.array_equal(so, sp)                 # 8x8 template bytes, sum, sum of squares
            f = ro["found"] == 1
            if subpix == 0:
                assert np.array_equal(ro["pos"], rp["pos"])                          # FindPatchCoarse: integer corner positions
            else:
                np.testing.assert_allclose(rp["pos"][f], ro["pos"][f], atol=1e-6, rtol=0)
            assert f.sum() > 20
        for override, mark in ((0.0, False), (1.0, False), (16.0, True)):
            mo, no = o.pose_update(override, mark)
            mp, np_ = p.pose_update(override, mark)
            assert np.array_equal(no, np_)
            np.testing.assert_allclose(mp, mo, atol=1e-9, rtol=0)
        for s in range(S):
            po, pp = o.get_points(s), p.get_points(s)
            assert np.array_equal(po["outliers"], pp["outliers"]) and np.array_equal(po["inliers"], pp["inliers"])


@pytest.mark.gpu
def test_cuda_units_leave_the_tracker_state_alone(product, scene):
    frames, poses, kfs, m = scene
    p = _tracker(product, kfs, m)
    p.set_state(0, pose12=poses[5])
    p.track_frames([frames[5]])
    before = p.get_state(0)
    p.patch_search_batch(np.asarray(poses[6]).reshape(1, 12), 10, 0)
    p.pose_update()
    after = p.get_state(0)
    assert list(before.se3_cam_from_world) == list(after.se3_cam_from_world)
    assert list(before.velocity) == list(after.velocity) and before.frame == after.frame


@pytest.mark.gpu
def test_cuda_units_errors(product, scene):
    from ptam_cg_b200.capi import PtamError
    frames, poses, kfs, m = scene
    p = _tracker(product, kfs, m)
    with pytest.raises(PtamError):
        p.pose_update()                       # nothing searched yet
    with pytest.raises(PtamError):
        p.patch_search_batch(np.asarray(poses[6]).reshape(1, 12), 10, 0)   # no current frame
