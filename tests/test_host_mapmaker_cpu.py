"""The MapMaker host mirror (ptam_cg_b200/host/MapMaker.h: BundleAdjustAll / BundleAdjustRecent / BundleAdjust,
reference src/MapMaker.cc:767-933) checked WITHOUT a GPU: mapmaker_check.cc is compiled against the CPU oracle,
which exports the product's ABI under another prefix (tests/orc_alias.h), and the map it leaves behind is compared
with what the reference's control flow must produce, computed in Python over the same ABI (tests/mapmaker_util.py).
The GPU run of the same binary against the CUDA library is tests/test_zz_host_mapmaker_gpu.py."""
import subprocess
from pathlib import Path

import pytest

import mapmaker_util as mu
from oracle.binding import oracle_lib

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "ptam_cg_b200" / "host"


@pytest.fixture(scope="module")
def orc_binary(tmp_path_factory):
    out = tmp_path_factory.mktemp("mm") / "mapmaker_check_orc"
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", str(HOST), "-include", str(ROOT / "tests" / "orc_alias.h"),
           str(HOST / "mapmaker_check.cc"), "-o", str(out), "-L", str(ROOT / "oracle"), "-loracle",
           "-Wl,-rpath," + str(ROOT / "oracle")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.mark.parametrize("mode", [0, 1], ids=["BundleAdjustAll", "BundleAdjustRecent"])
def test_mapmaker_mirror_follows_the_reference_control_flow(orc_binary, tmp_path, mode):
    g = mu.make_map()
    mu.write_map(g, tmp_path, mode, 20)
    r = subprocess.run([str(orc_binary), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = mu.read_map(tmp_path, len(g["cam_fixed"]), len(g["points"]))
    exp = mu.expected(oracle_lib(), g, mode, 20)
    assert exp["accepted"] > 0
    mu.compare(got, exp, tol=0.0)  # same library, same call order: bit for bit
    assert got["bad"].sum() + len(got["queue"]) + len(got["never"]) > 0  # the outlier bookkeeping was exercised


def test_recent_adjustment_needs_eight_keyframes(orc_binary, tmp_path):
    g = mu.make_map(n_cams=6, n_points=150, n_meas=600, seed=5)
    mu.write_map(g, tmp_path, 1, 20)
    r = subprocess.run([str(orc_binary), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = mu.read_map(tmp_path, 6, 150)
    import numpy as np
    assert np.array_equal(got["points"], g["points"]) and np.array_equal(got["cams"], g["cam_se3"])  # MapMaker.cc:789-792
    assert list(got["flags"][:2]) == [1, 1]


@pytest.fixture(scope="module")
def orc_libm_binary(tmp_path_factory):
    """mapmaker_check against liboracle_libm.so (platform atan: the variant pinned bit for bit against oracle/_ref)."""
    out = tmp_path_factory.mktemp("mm") / "mapmaker_check_orc_libm"
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", str(HOST), "-include", str(ROOT / "tests" / "orc_alias.h"),
           str(HOST / "mapmaker_check.cc"), "-o", str(out), "-L", str(ROOT / "oracle"), "-loracle_libm",
           "-Wl,-rpath," + str(ROOT / "oracle")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.mark.parametrize("pair", [(0, 30), (10, 25)], ids=lambda p: f"src{p[0]}-tgt{p[1]}")
def test_add_points_epipolar_matches_the_reference(orc_libm_binary, tmp_path, pair):
    """MapMaker::AddPointsEpipolar of the host mirror (device search + host triangulation and MapPoint construction)
    against the reference's own MapMaker::AddPointEpipolar compiled in place (oracle/_ref): same accepted
    candidates, same measurements, and the triangulated world position / pixel vectors of every new point
    (the reference asks an SVD<4> of the DLT matrix; the mirror runs a one-sided Jacobi SVD)."""
    import numpy as np
    from ptam_cg_b200 import synth
    from oracle.binding import ref_lib
    ref = ref_lib()
    if ref is None or not hasattr(ref.cdll, "ref_tracker_epipolar_last_points"):
        pytest.skip("oracle/_ref not built")
    W, H = 320, 240
    frames, poses = synth.render_sequence(W, H, 40)
    mu.write_epipolar_case(tmp_path, W, H, frames, poses, *pair)
    r = subprocess.run([str(orc_libm_binary), str(tmp_path), "epi"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = mu.read_epipolar_out(tmp_path)
    exp = mu.epipolar_expected(oracle_lib(libm_atan=True), W, H, frames, poses, *pair, ref=ref)
    assert np.array_equal(got["ncand"], exp["ncand"]) and np.array_equal(got["counts"], exp["counts"])
    assert got["counts"].sum() > 100
    assert np.array_equal(got["levels"], exp["levels"])
    assert np.array_equal(got["meas"], exp["meas"])                      # root position and sub-pixel target position
    np.testing.assert_allclose(got["points"][:, :3], exp["points"][:, :3], rtol=0, atol=1e-9)   # v3WorldPos
    np.testing.assert_allclose(got["points"][:, 3:], exp["points"][:, 3:], rtol=0, atol=1e-10)  # v3PixelRight_W / Down_W


def _refind_case(tmp_path, lib):
    """A 320x240 map from two source keyframes and a third frame to re-find the points in; a few points are
    marked as already measured / never to be retried in it.  Returns the expected per-point outcome from `lib`'s
    C ABI + the reference's bookkeeping (MapMaker.cc:943-1018)."""
    import numpy as np
    from ptam_cg_b200 import synth
    from ptam_cg_b200.capi import Tracker, PT_FOUND
    from oracle.binding import detect_with
    W, H = 320, 240
    frames, poses = synth.render_sequence(W, H, 10)
    cam = synth.AtanCamera(W, H)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle_lib(), W, H), cam, kf_indices=(0, 8), per_level=(150, 80, 40, 20))
    n = len(m["src_kf"])
    np.array([W, H, len(kfs), n], np.int32).tofile(tmp_path / "trk_dims.i32")
    np.ascontiguousarray(np.stack(kfs), np.uint8).tofile(tmp_path / "trk_kf.u8")
    for name, key, dt in (("trk_world.f64", "world_pos", np.float64), ("trk_right.f64", "pixel_right_w", np.float64),
                          ("trk_down.f64", "pixel_down_w", np.float64), ("trk_srckf.i32", "src_kf", np.int32),
                          ("trk_srclevel.i32", "src_level", np.int32), ("trk_center.i32", "ir_center", np.int32)):
        np.ascontiguousarray(m[key], dt).tofile(tmp_path / name)
    f = 4
    np.ascontiguousarray(frames[f], np.uint8).tofile(tmp_path / "rf_image.u8")
    np.ascontiguousarray(poses[f], np.float64).tofile(tmp_path / "rf_pose.f64")
    pre = np.zeros(n, np.int32)
    pre[::7] = 1
    pre[3::11] = 2
    pre.tofile(tmp_path / "rf_pre.i32")
    t = Tracker(lib, W, H, 1)
    for k in kfs:
        t.add_keyframe(k)
    t.set_map(0, m)
    t.refind_in_keyframes([frames[f]], [poses[f]])
    pts = t.get_points(0)
    found = (pts["flags"] & PT_FOUND) != 0
    exp = np.zeros((n, 5), np.int32)
    pos = np.zeros((n, 2))
    for i in range(n):
        if pre[i] == 1:
            exp[i] = (1, 0, 0, 0, 0); pos[i] = (-1.0, -1.0)       # the measurement it already had (SRC_TRACKER = 0)
        elif pre[i] == 2:
            exp[i] = (0, -1, -1, 0, 1)
        elif found[i]:
            exp[i] = (1, 1, pts["level"][i], int(pts["level"][i] > 0), 0)   # SRC_REFIND = 1
            pos[i] = pts["v2_found"][i]
        else:
            exp[i] = (0, -1, -1, 0, 1)
    return n, exp, pos, int((found & (pre == 0)).sum())


def test_refind_in_single_keyframe_bookkeeping(orc_binary, tmp_path):
    import numpy as np
    n, exp, pos, n_new = _refind_case(tmp_path, oracle_lib())
    r = subprocess.run([str(orc_binary), str(tmp_path), "refind"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "rf_out_points.i32", np.int32).reshape(n, 5)
    gpos = np.fromfile(tmp_path / "rf_out_pos.f64").reshape(n, 2)
    counts = np.fromfile(tmp_path / "rf_out_counts.i32", np.int32)
    assert list(counts[:2]) == [n_new, 0] and n_new > 100
    assert np.array_equal(got, exp)     # (the failure-queue round trip below restores every measurement it removed)
    assert np.array_equal(gpos, pos)
    # MapMaker::ReFindFromFailureQueue: every queued (keyframe, point) pair re-found, unchanged; queue emptied
    n_queued, n_second, n_same, left = counts[2:]
    assert n_queued > 10 and n_second == n_queued and n_same == n_queued and left == 0


def test_refind_newly_made_points_in_a_third_keyframe(orc_binary, tmp_path):
    """AddPointsEpipolar between frames 0 and 30, then a third keyframe joins the map and MapMaker::ReFindNewlyMade
    (MapMaker.cc:1046-1065) looks for the new points in it.  Expected: the same C ABI driven from Python on the map
    the mirror made (its points are checked against the reference in test_add_points_epipolar_matches_the_reference)."""
    import numpy as np
    from ptam_cg_b200 import synth
    from ptam_cg_b200.capi import Tracker, PT_FOUND
    W, H = 320, 240
    frames, poses = synth.render_sequence(W, H, 40)
    mu.write_epipolar_case(tmp_path, W, H, frames, poses, 0, 30)
    np.ascontiguousarray(frames[15], np.uint8).tofile(tmp_path / "epi_third.u8")
    np.ascontiguousarray(poses[15], np.float64).tofile(tmp_path / "epi_third_pose.f64")
    r = subprocess.run([str(orc_binary), str(tmp_path), "epi"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = mu.read_epipolar_out(tmp_path)
    n = len(got["points"])
    third = np.fromfile(tmp_path / "epi_out_third.i32", np.int32)
    pos3 = np.fromfile(tmp_path / "epi_out_third_pos.f64").reshape(n, 2)
    n_refound, queue_left = third[:2]
    per = third[2:].reshape(n, 4)   # measured in k3, level, never-retry in k3, GoodMeasCount
    # the same search through the C ABI: the new points as a map whose source keyframe is frame 0
    scale = (1 << got["levels"]).astype(np.float64)
    centre = np.round((got["meas"][:, :2] + 0.5) / scale[:, None] - 0.5).astype(np.int32)   # Candidate::irLevelPos
    m = dict(world_pos=got["points"][:, :3], pixel_right_w=got["points"][:, 3:6], pixel_down_w=got["points"][:, 6:9],
             src_kf=np.zeros(n, np.int32), src_level=got["levels"], ir_center=centre)
    t = Tracker(oracle_lib(), W, H, 1)
    t.add_keyframe(frames[0])
    t.set_map(0, m)
    t.refind_in_keyframes([frames[15]], [poses[15]])
    pts = t.get_points(0)
    found = (pts["flags"] & PT_FOUND) != 0
    assert n_refound == found.sum() and queue_left == 0 and found.sum() > 50
    assert np.array_equal(per[:, 0], found.astype(np.int32))
    assert np.array_equal(per[found, 1], pts["level"][found])
    assert np.array_equal(per[:, 2], (~found).astype(np.int32))
    assert np.array_equal(per[:, 3], 2 + found.astype(np.int32))       # source + target (+ the third keyframe)
    assert np.array_equal(pos3[found], pts["v2_found"][found])


def _add_keyframe_case(tmp_path, lib):
    """Inputs for `mapmaker_check <dir> addkf` and what MapMaker::AddKeyFrameFromTopOfQueue (MapMaker.cc:493-519) must
    produce, from `lib`'s C ABI and the reference's glue (ThinCandidates :413-441, ClosestKeyFrame :738-752,
    AddSomeMapPoints :449-458 on levels 3, 0, 1, 2)."""
    import numpy as np
    from ptam_cg_b200 import synth
    from ptam_cg_b200.capi import Tracker, PT_FOUND
    from oracle.binding import detect_with
    W, H = 320, 240
    frames, poses = synth.render_sequence(W, H, 12)
    cam = synth.AtanCamera(W, H)
    kf_idx = (0, 8)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle_lib(), W, H), cam, kf_indices=kf_idx, per_level=(150, 80, 40, 20))
    n, f = len(m["src_kf"]), 5   # nearer to keyframe 8 than to keyframe 0 (frame 4 would be a tie)
    depth, wiggle = (1.0, 0.3), 0.1
    np.array([W, H, len(kfs), n], np.int32).tofile(tmp_path / "trk_dims.i32")
    np.ascontiguousarray(np.stack(kfs), np.uint8).tofile(tmp_path / "trk_kf.u8")
    for name, key, dt in (("trk_world.f64", "world_pos", np.float64), ("trk_right.f64", "pixel_right_w", np.float64),
                          ("trk_down.f64", "pixel_down_w", np.float64), ("trk_srckf.i32", "src_kf", np.int32),
                          ("trk_srclevel.i32", "src_level", np.int32), ("trk_center.i32", "ir_center", np.int32)):
        np.ascontiguousarray(m[key], dt).tofile(tmp_path / name)
    np.ascontiguousarray([poses[i] for i in kf_idx], np.float64).tofile(tmp_path / "ak_kf_poses.f64")
    np.ascontiguousarray(frames[f], np.uint8).tofile(tmp_path / "rf_image.u8")
    np.ascontiguousarray(poses[f], np.float64).tofile(tmp_path / "rf_pose.f64")
    np.array([depth[0], depth[1], wiggle], np.float64).tofile(tmp_path / "ak_depth.f64")
    # what a search of every map point in the new frame finds; the tracker is given every other one of them
    t = Tracker(lib, W, H, 1)
    for k in kfs:
        t.add_keyframe(k)
    t.set_map(0, m)
    t.refind_in_keyframes([frames[f]], [poses[f]])
    pts = t.get_points(0)
    found = (pts["flags"] & PT_FOUND) != 0
    given = np.flatnonzero(found)[::2]
    np.ascontiguousarray(np.column_stack([given, pts["level"][given]]), np.int32).tofile(tmp_path / "ak_meas_idx.i32")
    np.ascontiguousarray(pts["v2_found"][given], np.float64).tofile(tmp_path / "ak_meas_pos.f64")
    meas = np.zeros((n, 5), np.int32)
    pos = np.zeros((n, 2))
    is_given = np.zeros(n, bool); is_given[given] = True
    for i in range(n):
        if found[i]:
            meas[i] = (1, 0 if is_given[i] else 1, pts["level"][i], 1, 0)   # SRC_TRACKER = 0, SRC_REFIND = 1
            pos[i] = pts["v2_found"][i]
        else:
            meas[i] = (0, -1, -1, 0, 1)
    # candidates of the new keyframe, the nearest keyframe, then levels 3, 0, 1, 2: thin, search, remember the roots
    t.make_keyframes([frames[f]])
    cands = [np.asarray(r[1], np.int32).reshape(-1, 2) for r in t.keyframe_rest(0, 70.0)]
    centre = lambda p: -synth.se3_from12(p)[0].T @ synth.se3_from12(p)[1]
    closest = int(np.argmin([np.linalg.norm(centre(poses[i]) - centre(poses[f])) for i in kf_idx]))
    t2 = Tracker(lib, W, H, 1)
    src_id = t2.add_keyframe(frames[f])
    t2.make_keyframes([frames[kf_idx[closest]]])
    busy_meas = [(int(pts["level"][i]), pts["v2_found"][i]) for i in np.flatnonzero(found)]
    rnd = lambda v: int(v + 0.5) if v > 0 else int(v - 0.5)
    thinned, new_counts = {}, {}
    for l in (3, 0, 1, 2):
        scale = 1 << l
        busy = [(rnd(p[0] / scale), rnd(p[1] / scale)) for lv, p in busy_meas if lv in (l, l + 1)]
        keep = [c for c in cands[l] if all((b[0] - c[0]) ** 2 + (b[1] - c[1]) ** 2 >= 100 for b in busy)]
        thinned[l] = np.array(keep, np.int32).reshape(-1, 2)
        if len(keep):
            fnd, _, _ = t2.epipolar_search(0, l, src_id, poses[f], depth[0], depth[1], poses[kf_idx[closest]], wiggle, thinned[l])
        else:
            fnd = np.zeros(0, np.int32)
        new_counts[l] = int(fnd.sum())
        for c in thinned[l][fnd != 0]:
            busy_meas.append((l, (np.asarray(c, np.float64) + 0.5) * scale - 0.5))
    return n, meas, pos, thinned, new_counts, closest


def _check_add_keyframe(tmp_path, n, meas, pos, thinned, new_counts, closest, tol):
    import numpy as np
    got_meas = np.fromfile(tmp_path / "ak_out_meas.i32", np.int32).reshape(n, 5)
    got_pos = np.fromfile(tmp_path / "ak_out_meas_pos.f64").reshape(n, 2)
    cand = np.fromfile(tmp_path / "ak_out_cand.i32", np.int32)
    new = np.fromfile(tmp_path / "ak_out_new.i32", np.int32)
    assert np.array_equal(got_meas, meas)
    np.testing.assert_allclose(got_pos, pos, rtol=0, atol=tol)
    o = 0
    for l in range(4):
        k = int(cand[o]); o += 1
        assert np.array_equal(cand[o:o + 2 * k].reshape(-1, 2), thinned[l]), f"thinned candidates of level {l}"
        o += 2 * k
    assert list(new[:4]) == [new_counts[l] for l in range(4)] and sum(new[:4]) > 50
    assert new[4] == closest


def test_add_keyframe_from_top_of_queue(orc_binary, tmp_path):
    case = _add_keyframe_case(tmp_path, oracle_lib())
    r = subprocess.run([str(orc_binary), str(tmp_path), "addkf"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    _check_add_keyframe(tmp_path, *case, tol=0.0)


def test_add_keyframe_from_top_of_queue_matches_the_reference_itself(orc_libm_binary, tmp_path):
    """The same hand-over through the reference's OWN MapMaker::AddKeyFrame + AddKeyFrameFromTopOfQueue, compiled in
    place (oracle/_ref, test hook ref_mapmaker_add_keyframe): the measurements the keyframe ends up with, the
    never-retry marks, the thinned candidate lists, the new points per level and the nearest keyframe must equal
    what the host mirror leaves behind (over the libm-atan oracle, which is pinned bit for bit against oracle/_ref)."""
    import ctypes as C
    import numpy as np
    from ptam_cg_b200 import synth
    from ptam_cg_b200.capi import Tracker
    from oracle.binding import ref_lib
    ref = ref_lib()
    if ref is None or not hasattr(ref.cdll, "ref_mapmaker_add_keyframe"):
        pytest.skip("oracle/_ref not built")
    n, meas, pos, thinned, new_counts, closest = _add_keyframe_case(tmp_path, oracle_lib(libm_atan=True))
    r = subprocess.run([str(orc_libm_binary), str(tmp_path), "addkf"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # ---- the reference, on the same files
    dims = np.fromfile(tmp_path / "trk_dims.i32", np.int32)
    W, H, nkf = int(dims[0]), int(dims[1]), int(dims[2])
    kfim = np.fromfile(tmp_path / "trk_kf.u8", np.uint8).reshape(nkf, H, W)
    kfpose = np.fromfile(tmp_path / "ak_kf_poses.f64").reshape(nkf, 12)
    m = dict(world_pos=np.fromfile(tmp_path / "trk_world.f64").reshape(n, 3), pixel_right_w=np.fromfile(tmp_path / "trk_right.f64").reshape(n, 3),
             pixel_down_w=np.fromfile(tmp_path / "trk_down.f64").reshape(n, 3), src_kf=np.fromfile(tmp_path / "trk_srckf.i32", np.int32),
             src_level=np.fromfile(tmp_path / "trk_srclevel.i32", np.int32), ir_center=np.fromfile(tmp_path / "trk_center.i32", np.int32).reshape(n, 2))
    t = Tracker(ref, W, H, 1)
    for k in range(nkf):
        t.add_keyframe(kfim[k])
        assert ref.cdll.ref_tracker_set_keyframe_pose(C.c_void_p(t.h), k, kfpose[k].ctypes.data_as(C.POINTER(C.c_double))) == 0
    t.set_map(0, m)
    image = np.fromfile(tmp_path / "rf_image.u8", np.uint8)
    pose = np.fromfile(tmp_path / "rf_pose.f64")
    depth = np.fromfile(tmp_path / "ak_depth.f64")
    tm_idx = np.fromfile(tmp_path / "ak_meas_idx.i32", np.int32)
    tm_pos = np.fromfile(tmp_path / "ak_meas_pos.f64")
    out_meas, out_pos = np.zeros((n, 5), np.int32), np.zeros((n, 2))
    out_cand, out_new = np.zeros(20000, np.int32), np.zeros(5, np.int32)
    fn = ref.cdll.ref_mapmaker_add_keyframe
    fn.restype = C.c_int
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    rc = fn(C.c_void_p(t.h), 0, image.ctypes.data_as(C.POINTER(C.c_uint8)), W, pose.ctypes.data_as(dp), C.c_double(depth[0]),
            C.c_double(depth[1]), C.c_double(depth[2]), len(tm_idx) // 2, tm_idx.ctypes.data_as(ip), tm_pos.ctypes.data_as(dp),
            out_meas.ctypes.data_as(ip), out_pos.ctypes.data_as(dp), out_cand.ctypes.data_as(ip), len(out_cand), out_new.ctypes.data_as(ip))
    assert rc > 0
    # ---- mirror == reference
    got_meas = np.fromfile(tmp_path / "ak_out_meas.i32", np.int32).reshape(n, 5)
    got_pos = np.fromfile(tmp_path / "ak_out_meas_pos.f64").reshape(n, 2)
    got_cand = np.fromfile(tmp_path / "ak_out_cand.i32", np.int32)
    got_new = np.fromfile(tmp_path / "ak_out_new.i32", np.int32)
    assert np.array_equal(got_meas, out_meas)
    assert np.array_equal(got_pos, out_pos)
    assert np.array_equal(got_cand, out_cand[:rc])
    assert np.array_equal(got_new, out_new) and got_new[:4].sum() > 50
    # and both equal the glue restated in Python
    _check_add_keyframe(tmp_path, n, meas, pos, thinned, new_counts, closest, tol=0.0)


@pytest.mark.parametrize("mode", [0, 1], ids=["BundleAdjustAll", "BundleAdjustRecent"])
def test_bundle_adjust_matches_the_reference_itself(orc_libm_binary, tmp_path, mode):
    """The reference's OWN MapMaker::BundleAdjustAll / BundleAdjustRecent over its own Bundle (oracle/_ref, test hook
    ref_mapmaker_bundle_adjust) on the same map: adjusted points and keyframes, bad flags, surviving measurements,
    failure queue (order included), never-retry sets and convergence flags must equal the host mirror's, bit for bit
    (the mirror runs over the libm-atan oracle, whose Bundle is pinned bit for bit against the reference's)."""
    import ctypes as C
    import numpy as np
    from ptam_cg_b200.capi import Tracker
    from oracle.binding import ref_lib
    ref = ref_lib()
    if ref is None or not hasattr(ref.cdll, "ref_mapmaker_bundle_adjust"):
        pytest.skip("oracle/_ref not built")
    g = mu.make_map()
    mu.write_map(g, tmp_path, mode, 20)
    r = subprocess.run([str(orc_libm_binary), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    Cn, Pn, Mn = len(g["cam_fixed"]), len(g["points"]), len(g["meas_cam"])
    got = mu.read_map(tmp_path, Cn, Pn)
    t = Tracker(ref, g["width"], g["height"], 1)   # lends its MapMaker and camera (640x480)
    f64 = lambda a: np.ascontiguousarray(a, np.float64)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    cams, fixed, pts = f64(g["cam_se3"]), i32(g["cam_fixed"]), f64(g["points"])
    mcam, mpt, uv, lvl, src = i32(g["meas_cam"]), i32(g["meas_point"]), f64(g["meas_uv"]), i32(g["meas_level"]), i32(g["meas_src"])
    o_pts, o_cams = np.zeros((Pn, 3)), np.zeros((Cn, 12))
    o_bad, o_nmeas = np.zeros(Pn, np.int32), np.zeros(Cn, np.int32)
    o_queue, o_never, o_nn, o_flags = np.zeros((Mn, 2), np.int32), np.zeros((Mn, 2), np.int32), np.zeros(1, np.int32), np.zeros(4, np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    D, I = (lambda a: a.ctypes.data_as(dp)), (lambda a: a.ctypes.data_as(ip))
    fn = ref.cdll.ref_mapmaker_bundle_adjust
    fn.restype = C.c_int
    nq = fn(C.c_void_p(t.h), 0, mode, 20, Cn, D(cams), I(fixed), Pn, D(pts), Mn, I(mcam), I(mpt), D(uv), I(lvl), I(src),
            D(o_pts), D(o_cams), I(o_bad), I(o_nmeas), I(o_queue), I(o_never), Mn, I(o_nn), I(o_flags))
    assert nq >= 0
    assert np.array_equal(got["points"], o_pts) and np.array_equal(got["cams"], o_cams)
    assert np.array_equal(got["bad"], o_bad) and np.array_equal(got["nmeas"], o_nmeas)
    assert np.array_equal(got["queue"], o_queue[:nq])
    assert np.array_equal(got["never"], o_never[:int(o_nn[0])])
    assert np.array_equal(got["flags"], o_flags)
    assert not np.array_equal(o_pts, pts)   # the adjustment did move the map


def _loop_case(tmp_path):
    """Inputs for `mapmaker_check <dir> loop`: a two-keyframe map, a short sequence, the frame that becomes a keyframe."""
    import numpy as np
    from ptam_cg_b200 import synth
    from ptam_cg_b200.capi import Tracker
    from oracle.binding import detect_with
    W, H, first, nfr, hand_over = 320, 240, 2, 8, 3
    frames, poses = synth.render_sequence(W, H, 16)
    cam = synth.AtanCamera(W, H)
    kf_idx = (0, 12)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle_lib(), W, H), cam, kf_indices=kf_idx, per_level=(150, 80, 40, 20))
    n = len(m["src_kf"])
    np.array([W, H, len(kfs), n, nfr, hand_over], np.int32).tofile(tmp_path / "trk_dims.i32")
    np.ascontiguousarray(np.stack(kfs), np.uint8).tofile(tmp_path / "trk_kf.u8")
    np.ascontiguousarray([poses[i] for i in kf_idx], np.float64).tofile(tmp_path / "ak_kf_poses.f64")
    np.ascontiguousarray(frames[first:first + nfr], np.uint8).tofile(tmp_path / "trk_frames.u8")
    for name, key, dt in (("trk_world.f64", "world_pos", np.float64), ("trk_right.f64", "pixel_right_w", np.float64),
                          ("trk_down.f64", "pixel_down_w", np.float64), ("trk_srckf.i32", "src_kf", np.int32),
                          ("trk_srclevel.i32", "src_level", np.int32), ("trk_center.i32", "ir_center", np.int32)):
        np.ascontiguousarray(m[key], dt).tofile(tmp_path / name)
    np.ascontiguousarray(synth.perturb_pose(poses[first], np.random.default_rng(3)), np.float64).tofile(tmp_path / "trk_pose0.f64")
    return n, nfr, hand_over, [poses[first + i] for i in range(nfr)], m


def _check_loop(tmp_path, n, nfr, hand_over, truth, m):
    import numpy as np
    from ptam_cg_b200 import synth
    poses = np.fromfile(tmp_path / "loop_out_poses.f64").reshape(nfr + 1, 12)
    found = np.fromfile(tmp_path / "loop_out_found.i32", np.int32)
    (new_pts, kf_meas, refound, n_points, n_kfs, reset, converged, n_bad, queue, passes, dangling, n_live,
     n_trash) = np.fromfile(tmp_path / "loop_out_info.i32", np.int32)
    for f in range(nfr):   # the tracker follows the ground truth throughout, before and after the map grew
        assert np.abs(poses[f][9:] - synth.se3_from12(truth[f])[1]).max() < 5e-3, f
    # MapMaker::RunOnce until the thread's priority list is empty: converged, bad points gone with their measurements
    assert converged == 1 and passes < 12 and dangling == 0 and n_live + n_trash == n_points and n_trash == n_bad
    # ... and the tracker re-tracks the last frame on the adjusted, cleaned-up map (it was re-uploaded)
    assert np.abs(poses[nfr][9:] - synth.se3_from12(truth[nfr - 1])[1]).max() < 5e-3 and found[nfr] > 0.5 * found[nfr - 1]
    found = found[:nfr]
    assert new_pts > 50 and n_points == n + new_pts and n_kfs == 3
    assert kf_meas >= found[hand_over] + new_pts          # the tracker's measurements + re-found ones + the new points' roots
    assert found[hand_over + 1:].min() > found[:hand_over + 1].max() + new_pts // 2   # the new points are tracked at once
    # (with three keyframes most points have two measurements, and an outlier measurement of such a point makes it
    # bad, MapMaker.cc:919-920: a quarter of this young map goes that way under the Tukey estimator)
    assert refound == new_pts and reset == 0 and queue == 0 and n_bad < n_points // 2   # the new-point queue was worked off
    before = np.fromfile(tmp_path / "loop_out_before_ba.f64").reshape(n_points, 3)
    after = np.fromfile(tmp_path / "loop_out_after_ba.f64").reshape(n_points, 3)
    assert np.isfinite(after).all() and not np.array_equal(before, after)      # the adjustment ran and moved the map ...
    assert np.abs(after - before).max() < 0.05                                  # ... a little: it was consistent already
    # the planar scene: every point, old and new, sits on z = 0 (the new ones were triangulated, not intersected)
    # (5.6e-3 of the scene depth on this young three-keyframe map, CPU oracle and CUDA library alike)
    assert np.abs(after[:, 2]).mean() < 1e-2
    assert np.array_equal(np.fromfile(tmp_path / "loop_out_kf0.f64"), np.fromfile(tmp_path / "ak_kf_poses.f64")[:12])  # gauge


def test_tracker_and_mapmaker_loop(orc_binary, tmp_path):
    """Both mirrors in the reference's loop: track, hand a keyframe to the map maker, track on with the enlarged map,
    re-find the new points, bundle-adjust everything (see run_loop in host/mapmaker_check.cc)."""
    case = _loop_case(tmp_path)
    r = subprocess.run([str(orc_binary), str(tmp_path), "loop"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    _check_loop(tmp_path, *case)


@pytest.mark.parametrize("params", [(0.3, 0.13, 0.1), (0.999, 0.0, 0.0005), (0.999, 0.0, 0.5)],
                         ids=["defaults-keyframe-drops", "far-from-keyframes-lost", "inconclusive-but-near"])
def test_tracker_consults_the_map_maker_like_the_reference(orc_libm_binary, tmp_path, params):
    """With a map maker attached the Tracker mirror also runs the last branch of AssessTrackingQuality
    (Tracker.cc:1094-1099: far from every keyframe = lost) and the keyframe hand-over heuristic (Tracker.cc:146-166).
    Against the reference's OWN Tracker + MapMaker (oracle/_ref; hook ref_tracker_mapmaker_ctl sets mdWiggleScale and
    reads QueueSize): tracking quality, lost-frame counter and queue length after every frame must be equal.
    Parameter sets: the defaults (quality good: keyframes are dropped at frames 1, 22, 43 until three wait); a
    quality threshold nothing reaches with a tiny / a large wiggle scale (the distance branch decides every frame)."""
    import ctypes as C
    import numpy as np
    from ptam_cg_b200 import synth
    from ptam_cg_b200.capi import Tracker
    from oracle.binding import detect_with, ref_lib
    ref = ref_lib()
    if ref is None or not hasattr(ref.cdll, "ref_tracker_mapmaker_ctl"):
        pytest.skip("oracle/_ref not built")
    good, lost, wiggle = params
    W, H, first, nfr, frames, poses, kf_idx, kfs, m, pose0 = _heuristics_case(tmp_path, params)
    r = subprocess.run([str(orc_libm_binary), str(tmp_path), "heur"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "hq_out.i32", np.int32).reshape(nfr, 4)
    _check_heuristics_against_reference(ref, params, W, H, first, nfr, frames, poses, kf_idx, kfs, m, pose0, got)


def _heuristics_case(tmp_path, params):
    import numpy as np
    from ptam_cg_b200 import synth
    from ptam_cg_b200.capi import Tracker
    from oracle.binding import detect_with
    good, lost, wiggle = params
    W, H, first, nfr = 320, 240, 1, 46
    frames, poses = synth.render_sequence(W, H, 48)
    cam = synth.AtanCamera(W, H)
    kf_idx = (0, 40)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle_lib(), W, H), cam, kf_indices=kf_idx, per_level=(150, 80, 40, 20))
    n = len(m["src_kf"])
    np.array([W, H, len(kfs), n, nfr], np.int32).tofile(tmp_path / "trk_dims.i32")
    np.ascontiguousarray(np.stack(kfs), np.uint8).tofile(tmp_path / "trk_kf.u8")
    np.ascontiguousarray([poses[i] for i in kf_idx], np.float64).tofile(tmp_path / "ak_kf_poses.f64")
    np.ascontiguousarray(frames[first:first + nfr], np.uint8).tofile(tmp_path / "trk_frames.u8")
    for name, key, dt in (("trk_world.f64", "world_pos", np.float64), ("trk_right.f64", "pixel_right_w", np.float64),
                          ("trk_down.f64", "pixel_down_w", np.float64), ("trk_srckf.i32", "src_kf", np.int32),
                          ("trk_srclevel.i32", "src_level", np.int32), ("trk_center.i32", "ir_center", np.int32)):
        np.ascontiguousarray(m[key], dt).tofile(tmp_path / name)
    pose0 = synth.perturb_pose(poses[first], np.random.default_rng(3))
    np.ascontiguousarray(pose0, np.float64).tofile(tmp_path / "trk_pose0.f64")
    np.array([good, lost, wiggle], np.float64).tofile(tmp_path / "hq_params.f64")
    return W, H, first, nfr, frames, poses, kf_idx, kfs, m, pose0


def _check_heuristics_against_reference(ref, params, W, H, first, nfr, frames, poses, kf_idx, kfs, m, pose0, got):
    import ctypes as C
    import numpy as np
    from ptam_cg_b200.capi import Tracker
    good, lost, wiggle = params
    # ---- the reference's own tracker and map maker
    t = Tracker(ref, W, H, 1, quality_good=good, quality_lost=lost)
    ctl = ref.cdll.ref_tracker_mapmaker_ctl
    ctl.restype = C.c_int
    for k, im in enumerate(kfs):
        t.add_keyframe(im)
        pk = np.ascontiguousarray(poses[kf_idx[k]], np.float64)
        assert ref.cdll.ref_tracker_set_keyframe_pose(C.c_void_p(t.h), k, pk.ctypes.data_as(C.POINTER(C.c_double))) == 0
    t.set_map(0, m)
    t.set_state(0, pose12=pose0, velocity=np.zeros(6), msd=0.0)
    assert ctl(C.c_void_p(t.h), 0, C.c_double(wiggle), 2) == 0   # keep the queue: no map-maker thread drains it here
    exp = []
    for f in range(nfr):
        t.track_frames([frames[first + f]])
        st = t.get_state(0)
        exp.append((st.tracking_quality, st.lost_frames, ctl(C.c_void_p(t.h), 0, C.c_double(-1.0), 0)))
    exp = np.array(exp, np.int32)
    ctl(C.c_void_p(t.h), 0, C.c_double(1e30), 3)
    assert np.array_equal(got[:, :3], exp), (got[:8], exp[:8])
    if good < 0.5:
        assert list(np.flatnonzero(np.diff(np.r_[0, got[:, 2]]))) == [0, 21, 42] and got[-1, 2] == 3   # frames 1, 22, 43
    elif wiggle < 0.01:
        assert got[:, 3].sum() > 0 and got[:, 1].max() >= 1      # the distance branch fired and declared the tracker lost
    else:
        assert got[:, 3].sum() > 0 and (got[:, 0] == 2).all() and (got[:, 1] == 0).all() and got[0, 2] == 1
