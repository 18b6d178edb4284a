"""CPU-only checks that pin the oracle itself: the reference has no tests, so every
known-answer vector here is derived from the published definitions of the primitives."""
import ctypes as C

import numpy as np
import pytest

from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Tracker, CAMERA_PARAMS
from oracle.binding import oracle_lib, fast10_bruteforce

RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
        (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def kf(oracle, im):
    h, w = im.shape
    t = Tracker(oracle, w, h)
    t.make_keyframes([im])
    return t


# ---------------------------------------------------------------- FAST-10 / halfSample / LUT
@pytest.mark.parametrize("arc,expect", [(9, False), (10, True), (11, True), (16, True)])
@pytest.mark.parametrize("bright", [True, False])
def test_fast_arc_length_every_rotation(oracle, arc, expect, bright):
    for rot in range(16):
        im = np.full((64, 64), 100, np.uint8)
        for k in range(arc):
            dx, dy = RING[(rot + k) % 16]
            im[30 + dy, 30 + dx] = 111 if bright else 89  # threshold 10: strictly beyond p +- t
        _, xy, _ = kf(oracle, im).get_level(0, 0)
        got = any((x, y) == (30, 30) for x, y in xy)
        assert got == expect, (arc, rot, bright)


def test_fast_threshold_is_strict_and_border_excluded(oracle):
    im = np.full((64, 64), 100, np.uint8)
    for k in range(12):
        dx, dy = RING[k]
        im[30 + dy, 30 + dx] = 110  # == p + t: not brighter
    _, xy, _ = kf(oracle, im).get_level(0, 0)
    assert not any((x, y) == (30, 30) for x, y in xy)
    rng = np.random.default_rng(0)
    im = rng.integers(0, 256, (70, 90), dtype=np.uint8)
    _, xy, lut = kf(oracle, im).get_level(0, 0)
    assert len(xy) > 50
    assert xy[:, 0].min() >= 3 and xy[:, 1].min() >= 3 and xy[:, 0].max() < 90 - 3 and xy[:, 1].max() < 70 - 3
    # raster order and LUT definition: LUT[y] = number of corners in rows < y
    key = xy[:, 1].astype(np.int64) * 1000 + xy[:, 0]
    assert np.all(np.diff(key) > 0)
    assert np.array_equal(lut, [int((xy[:, 1] < y).sum()) for y in range(70)])


@pytest.mark.parametrize("seed", range(4))
def test_fast_matches_bruteforce_definition(oracle, seed):
    rng = np.random.default_rng(seed)
    tex = synth.make_texture(seed=seed + 1, size=256, n_rects=150)
    for im in (rng.integers(0, 256, (80, 100), dtype=np.uint8), tex[:120, :160].copy(),
               (rng.integers(0, 4, (64, 64)) * 80).astype(np.uint8)):
        t = kf(oracle, im)
        for l, thr in enumerate((10, 15, 15, 10)):
            pix, xy, _ = t.get_level(0, l)
            assert np.array_equal(xy, fast10_bruteforce(pix, thr))


def test_fast10_subset_of_cv2_fast9(oracle):
    cv2 = pytest.importorskip("cv2")
    tex = synth.make_texture(seed=5, size=512, n_rects=500)
    im = tex[:240, :320].copy()
    _, xy, _ = kf(oracle, im).get_level(0, 0)
    det = cv2.FastFeatureDetector_create(threshold=10, nonmaxSuppression=False, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    nine = {(int(k.pt[0]), int(k.pt[1])) for k in det.detect(im, None)}
    ours = {(int(x), int(y)) for x, y in xy}
    assert len(ours) > 100 and ours <= nine


def test_half_sample_truncates(oracle):
    im = np.zeros((64, 64), np.uint8)
    im[0:2, 0:2] = [[255, 255], [255, 254]]
    im[0:2, 2:4] = [[0, 0], [0, 3]]
    im[2:4, 0:2] = [[1, 2], [3, 5]]
    t = kf(oracle, im)
    l1 = t.get_level(0, 1)[0]
    assert l1.shape == (32, 32) and l1[0, 0] == 254 and l1[0, 1] == 0 and l1[1, 0] == 2
    im = np.arange(67 * 65, dtype=np.uint32).reshape(65, 67).astype(np.uint8)
    t = kf(oracle, np.ascontiguousarray(im))
    assert [t.level_size(l) for l in range(4)] == [(67, 65), (33, 32), (16, 16), (8, 8)]
    l1 = t.get_level(0, 1)[0]
    ref = (im[0:64:2, 0:66:2].astype(int) + im[0:64:2, 1:66:2] + im[1:64:2, 0:66:2] + im[1:64:2, 1:66:2]) // 4
    assert np.array_equal(l1, ref)


# ---------------------------------------------------------------- ZMSSD
def zmssd(oracle, im, x, y, tmpl):
    im = np.ascontiguousarray(im, np.uint8)
    tmpl = np.ascontiguousarray(tmpl, np.uint8)
    p8 = C.POINTER(C.c_uint8)
    return oracle.cdll.orc_zmssd(im.ctypes.data_as(p8), im.shape[1], im.shape[0], x, y, tmpl.ctypes.data_as(p8))


def test_zmssd_known_answers(oracle):
    rng = np.random.default_rng(1)
    im = rng.integers(40, 200, (40, 40), dtype=np.uint8)
    patch = im[10 - 4:10 + 4, 20 - 4:20 + 4].copy()
    assert zmssd(oracle, im, 20, 10, patch) == 0                      # identical patch
    assert zmssd(oracle, im, 20, 10, patch + 17) == 0                 # zero-mean: constant offset
    assert zmssd(oracle, im, 3, 10, patch) == 32001                   # not 4 px inside: max + 1
    assert zmssd(oracle, im, 20, 36, patch) == 32001
    # hand formula incl. C++ truncating division of a non-positive numerator
    t = rng.integers(0, 256, (8, 8), dtype=np.uint8)
    i = im[16:24, 6:14].astype(np.int64)
    tt = t.astype(np.int64)
    SA, SB = int(tt.sum()), int(i.sum())
    num = 2 * SA * SB - SA * SA - SB * SB
    assert num < 0 and num % 64 != 0                                   # truncation matters here
    expect = int(np.trunc(num / 64)) + int((i * i).sum()) + int((tt * tt).sum()) - 2 * int((i * tt).sum())
    assert zmssd(oracle, im, 10, 20, t) == expect
    assert expect != (num >> 6) + int((i * i).sum()) + int((tt * tt).sum()) - 2 * int((i * tt).sum())


# ---------------------------------------------------------------- camera, atan, SE3
def test_spec_atan_within_one_ulp_of_libm(oracle):
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-3, 3, 20000), rng.uniform(-1e-3, 1e-3, 2000), 10 ** rng.uniform(-10, 6, 2000),
                         [0.0, 0.4375, 0.6875, 1.1875, 2.4375, 1e-9, 1e20, -1e20]])
    got = np.array([oracle.cdll.orc_atan(float(x)) for x in xs])
    ref = np.arctan(xs)
    ulp = np.spacing(np.abs(ref))
    assert np.all(np.abs(got - ref) <= ulp)
    assert (got == ref).mean() > 0.9


def test_camera_roundtrip_and_derivs(oracle):
    rng = np.random.default_rng(2)
    cp = np.ascontiguousarray(CAMERA_PARAMS)
    for _ in range(200):
        cam = rng.uniform(-0.4, 0.4, 2)
        im, d, back = np.zeros(2), np.zeros(4), np.zeros(2)
        inv = C.c_int()
        oracle.cdll.orc_cam_project(dp(cp), 640, 480, dp(cam), dp(im), dp(d), C.byref(inv))
        oracle.cdll.orc_cam_unproject(dp(cp), 640, 480, dp(im), dp(back))
        np.testing.assert_allclose(back, cam, atol=1e-12)
        h = 1e-6
        num = np.zeros((2, 2))
        for k in range(2):
            a, b = cam.copy(), cam.copy()
            a[k] += h; b[k] -= h
            ia, ib = np.zeros(2), np.zeros(2)
            oracle.cdll.orc_cam_project(dp(cp), 640, 480, dp(a), dp(ia), None, None)
            oracle.cdll.orc_cam_project(dp(cp), 640, 480, dp(b), dp(ib), None, None)
            num[:, k] = (ia - ib) / (2 * h)
        if np.hypot(*cam) > 0.011:  # below r = 0.01 the reference zeroes the radial term by design
            np.testing.assert_allclose(d.reshape(2, 2), num, rtol=1e-6, atol=1e-6)
    # w = 0: plain pinhole
    pin = np.array([1.0, 1.3, 0.5, 0.5, 0.0])
    cam = np.array([0.2, -0.1]); im = np.zeros(2); d = np.zeros(4)
    oracle.cdll.orc_cam_project(dp(pin), 640, 480, dp(cam), dp(im), dp(d), None)
    np.testing.assert_allclose(im, [640 * 0.5 - 0.5 + 640 * 0.2, 480 * 0.5 - 0.5 - 480 * 1.3 * 0.1])
    np.testing.assert_allclose(d, [640.0, 0, 0, 480 * 1.3])


def test_se3_exp_ln(oracle):
    rng = np.random.default_rng(3)
    for scale in (1e-6, 1e-4, 5e-4, 1e-2, 0.3, 2.0):
        for _ in range(20):
            mu = rng.normal(0, scale, 6)
            out, back = np.zeros(12), np.zeros(6)
            oracle.cdll.orc_se3_exp(dp(mu), dp(out))
            R, t = synth.se3_exp(mu)
            np.testing.assert_allclose(out[:9].reshape(3, 3), R, atol=1e-12)
            np.testing.assert_allclose(out[9:], t, atol=1e-12)
            np.testing.assert_allclose(out[:9].reshape(3, 3) @ out[:9].reshape(3, 3).T, np.eye(3), atol=1e-12)
            oracle.cdll.orc_se3_ln(dp(out), dp(back))
            again = np.zeros(12)
            oracle.cdll.orc_se3_exp(dp(back), dp(again))
            np.testing.assert_allclose(again, out, atol=1e-9)          # exp(ln(T)) == T, also beyond pi
            if np.linalg.norm(mu[3:]) < 3.0:
                np.testing.assert_allclose(back, mu, atol=1e-9 * max(1, scale), rtol=1e-7)


# ---------------------------------------------------------------- tracker level
@pytest.fixture(scope="module")
def small_scene(oracle):
    from oracle.binding import detect_with
    W, H = 320, 240
    frames, poses = synth.render_sequence(W, H, 10)
    cam = synth.AtanCamera(W, H)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle, W, H), cam, kf_indices=(0, 5),
                             per_level=(200, 100, 50, 25))
    return frames, poses, kfs, m


def _tracker(oracle, small_scene, **prm):
    frames, poses, kfs, m = small_scene
    t = Tracker(oracle, 320, 240, 1, **prm)
    for k in kfs:
        t.add_keyframe(k)
    t.set_map(0, m)
    return t


def test_planted_template_found_at_source_corner(oracle, small_scene):
    """Tracking the source keyframe itself from its true pose: identity warp, every level-0 point
    must be found exactly at its own corner with ZMSSD 0."""
    frames, poses, kfs, m = small_scene
    t = _tracker(oracle, small_scene, disable_coarse=1)
    t.set_state(0, pose12=poses[0])
    r = t.track_frames([frames[0]])[0]
    pts = t.get_points(0)
    own = (m["src_kf"] == 0) & ((pts["flags"] & 8) != 0) & (pts["level"] == m["src_level"])
    assert own.sum() > 50
    sc = (1 << m["src_level"][own]).astype(float)
    expect = (m["ir_center"][own] + 0.5) * sc[:, None] - 0.5
    sub = (pts["flags"][own] & 16) != 0
    assert np.array_equal(pts["v2_found"][own][~sub], expect[~sub])
    np.testing.assert_allclose(pts["v2_found"][own][sub], expect[sub], atol=0.3)
    tm, _ = t.get_templates(0)
    src = kfs[0]
    i = np.flatnonzero(own & (m["src_level"] == 0))[0]
    x, y = m["ir_center"][i]
    assert np.abs(tm[i].reshape(8, 8).astype(int) - src[y - 4:y + 4, x - 4:x + 4].astype(int)).max() <= 1
    np.testing.assert_allclose(np.array(r.se3_cam_from_world), poses[0], atol=2e-3)


def test_pose_update_recovers_known_perturbation(oracle, small_scene):
    frames, poses, kfs, m = small_scene
    t = _tracker(oracle, small_scene)
    rng = np.random.default_rng(5)
    t.set_state(0, pose12=synth.perturb_pose(poses[3], rng, sigma=0.004))
    r = t.track_frames([frames[3]])[0]
    assert sum(r.meas_found) > 100
    np.testing.assert_allclose(np.array(r.se3_cam_from_world)[9:], poses[3][9:], atol=3e-3)
    np.testing.assert_allclose(np.array(r.se3_cam_from_world)[:9], poses[3][:9], atol=3e-3)
    assert r.tracking_quality == 2


def test_libm_atan_variant_gives_same_integer_outputs(small_scene):
    """Quantifies the specified-atan choice (oracle_math.h): with glibc's atan instead, templates,
    levels and found flags are unchanged on this scene (they may differ by rare last-bit flips)."""
    frames, poses, kfs, m = small_scene
    out = []
    for libm in (False, True):
        t = _tracker(oracle_lib(libm_atan=libm), small_scene)
        t.set_state(0, pose12=synth.perturb_pose(poses[4], np.random.default_rng(1)))
        t.track_frames([frames[4]])
        out.append((t.get_points(0), t.get_templates(0)[0]))
    (pa, ta), (pb, tb) = out
    assert np.array_equal(pa["level"], pb["level"])
    assert (ta != tb).mean() < 1e-3
    assert (pa["flags"] != pb["flags"]).mean() < 0.01


def test_sbi_rotation_estimator_known_answers(oracle):
    """SmallBlurryImage + CalcSBIRotation (ImageProcess.cc:279-494): identical frames
    give no rotation; an in-plane camera roll of theta is recovered about the optical axis with the
    right sign; the small image is ~zero-mean and has the documented size."""
    W, H = 640, 480
    tex = synth.make_texture()
    cam = synth.AtanCamera(W, H)
    _, poses = synth.render_sequence(W, H, 2)
    base = poses[0]
    theta = 0.03
    rolled = synth.se3_to12(*synth.se3_mul(synth.se3_exp([0, 0, 0, 0, 0, theta]), synth.se3_from12(base)))
    f0, f1 = synth.render_frame(tex, cam, base), synth.render_frame(tex, cam, rolled)
    t = Tracker(oracle, W, H, 1)
    t.track_frames([f0])
    tm, rot, score = t.get_sbi(0)
    assert tm.shape == (30, 40) and abs(float(tm.mean())) < 0.5
    assert np.allclose(rot, 0, atol=1e-9) and score < 1e-6          # first frame: both small images are the same
    t.track_frames([f0])
    _, rot, score = t.get_sbi(0)
    assert np.allclose(rot, 0, atol=1e-9)
    t.track_frames([f1])
    _, rot, _ = t.get_sbi(0)
    assert abs(rot[2] - theta) < 0.3 * theta and abs(rot[0]) < 0.01 and abs(rot[1]) < 0.01, rot
    # with the estimator the predicted pose (no map: TrackMap finds nothing) moves towards the roll
    st = t.get_state(0)
    t2 = Tracker(oracle, W, H, 1, use_rotation_estimator=0)
    for f in (f0, f0, f1):
        t2.track_frames([f])
    assert np.allclose(np.array(t2.get_state(0).se3_cam_from_world), np.array(Tracker(oracle, W, H, 1).get_state(0).se3_cam_from_world))
    assert not np.allclose(np.array(st.se3_cam_from_world)[:9], np.eye(3).reshape(9), atol=1e-4)


def test_keyframe_rest_known_answers(oracle):
    """fast_nonmax + Shi-Tomasi candidates: max corners are a subset of the FAST
    corners in raster order, no kept corner has a stronger FAST corner among its 8 neighbours, an
    isolated corner survives, candidates are max corners >= 10 px inside with score above the threshold,
    and the Shi-Tomasi score equals a numpy evaluation of ImageProcess.cc:20-47."""
    W, H = 320, 240
    tex = synth.make_texture(seed=5)
    im = tex[300:300 + H, 500:500 + W].copy()
    t = Tracker(oracle, W, H, 1)
    t.make_keyframes([im])
    rest = t.keyframe_rest(0, 70.0)
    for l in range(4):
        pix, corners, _ = t.get_level(0, l)
        mx, cx, cs = rest[l]
        cset = {tuple(c) for c in corners.tolist()}
        assert all(tuple(c) in cset for c in mx.tolist())
        order = [y * 100000 + x for x, y in mx.tolist()]
        assert order == sorted(order)
        mset = {tuple(c) for c in mx.tolist()}
        assert all(tuple(c) in mset for c in cx.tolist())
        hh, ww = pix.shape
        for (x, y), sc in zip(cx.tolist(), cs.tolist()):
            assert 10 <= x < ww - 10 and 10 <= y < hh - 10 and sc > 70.0
            win = pix.astype(np.float64)
            gx = win[y - 3:y + 4, x - 2:x + 5] - win[y - 3:y + 4, x - 4:x + 3]
            gy = win[y - 2:y + 5, x - 3:x + 4] - win[y - 4:y + 3, x - 3:x + 4]
            xx, yy, xy = (gx * gx).sum() / 98.0, (gy * gy).sum() / 98.0, (gx * gy).sum() / 98.0
            ref = 0.5 * (xx + yy - np.sqrt((xx + yy) ** 2 - 4 * (xx * yy - xy * xy)))
            assert abs(ref - sc) <= 1e-9 * max(1.0, abs(ref))
        assert len(mx) <= len(corners) and (len(corners) == 0 or len(mx) > 0)
    # an isolated bright square corner on a flat background is its own maximum
    flat = np.full((64, 64), 50, np.uint8)
    flat[30:, 30:] = 200
    t2 = Tracker(oracle, 64, 64, 1)
    t2.make_keyframes([flat])
    c0 = t2.get_level(0, 0)[1]
    m0 = t2.keyframe_rest(0, 0.0)[0][0]
    assert 0 < len(m0) <= len(c0)
