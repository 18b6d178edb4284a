"""Shared by the CPU and GPU tests of the MapMaker host mirror (ptam_cg_b200/host/MapMaker.h): writes a
synthetic map as raw arrays for mapmaker_check, and computes what MapMaker::BundleAdjustAll / BundleAdjustRecent
(reference src/MapMaker.cc:767-933) must leave in the map, by driving the same C ABI from Python in the
reference's own order (cameras: adjust set then fixed set, each in pointer = index order; points in index
order; measurements keyframe by keyframe in vpKeyFrames order, points ascending inside a keyframe)."""
import numpy as np

from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Bundle

SRC_TRACKER, SRC_REFIND, SRC_ROOT, SRC_TRAIL, SRC_EPIPOLAR = range(5)


def make_map(n_cams=12, n_points=400, n_meas=1600, seed=31):
    g = synth.make_ba_graph(n_cams, n_points, n_meas, seed=seed)
    rng = np.random.default_rng(seed)
    level = np.round(np.log2(np.sqrt(g["meas_sigma_sq"]))).astype(np.int32)
    src = rng.choice([SRC_TRACKER, SRC_REFIND, SRC_TRAIL, SRC_EPIPOLAR], size=len(level)).astype(np.int32)
    first = {}
    for i, p in enumerate(g["meas_point"]):  # every point's first measurement is its root measurement
        first.setdefault(int(p), i)
    src[list(first.values())] = SRC_ROOT
    g = dict(g)
    g["meas_level"], g["meas_src"] = level, src
    return g


def write_map(g, d, mode, max_iterations):
    np.ascontiguousarray(g["cam_se3"], np.float64).tofile(d / "mm_cams.f64")
    np.ascontiguousarray(g["cam_fixed"], np.int32).tofile(d / "mm_fixed.i32")
    np.ascontiguousarray(g["points"], np.float64).tofile(d / "mm_pts.f64")
    np.ascontiguousarray(g["meas_cam"], np.int32).tofile(d / "mm_mcam.i32")
    np.ascontiguousarray(g["meas_point"], np.int32).tofile(d / "mm_mpt.i32")
    np.ascontiguousarray(g["meas_uv"], np.float64).tofile(d / "mm_uv.f64")
    np.ascontiguousarray(g["meas_level"], np.int32).tofile(d / "mm_level.i32")
    np.ascontiguousarray(g["meas_src"], np.int32).tofile(d / "mm_src.i32")
    np.array([mode, max_iterations], np.int32).tofile(d / "mm_mode.i32")


def read_map(d, n_cams, n_points):
    out = dict(points=np.fromfile(d / "mm_out_pts.f64").reshape(n_points, 3), cams=np.fromfile(d / "mm_out_cams.f64").reshape(n_cams, 12),
               bad=np.fromfile(d / "mm_out_bad.i32", np.int32), nmeas=np.fromfile(d / "mm_out_nmeas.i32", np.int32),
               queue=np.fromfile(d / "mm_out_queue.i32", np.int32).reshape(-1, 2), never=np.fromfile(d / "mm_out_never.i32", np.int32).reshape(-1, 2),
               flags=np.fromfile(d / "mm_out_flags.i32", np.int32))
    return out


def _centre(se3):
    R, t = synth.se3_from12(se3)
    return -R.T @ t


def expected(lib, g, mode, max_iterations):
    C, P = len(g["cam_fixed"]), len(g["points"])
    meas_of_kf = {c: {} for c in range(C)}   # keyframe -> {point: measurement index}
    kfs_of_pt = {p: set() for p in range(P)}
    for i, (c, p) in enumerate(zip(g["meas_cam"], g["meas_point"])):
        meas_of_kf[int(c)][int(p)] = i
        kfs_of_pt[int(p)].add(int(c))
    fixed = [bool(f) for f in g["cam_fixed"]]
    flags = dict(full=True, recent=True)
    if mode == 0:
        adjust = [c for c in range(C) if not fixed[c]]
        fixed_set = [c for c in range(C) if fixed[c]]
        points = list(range(P))
    else:
        if C < 8:
            return None
        newest = C - 1
        centres = [_centre(g["cam_se3"][c]) for c in range(C)]
        order = sorted((float(np.sqrt(((centres[c] - centres[newest]) ** 2).sum())), c) for c in range(C) if c != newest)
        adjust = sorted({newest} | {c for _, c in order[:4] if not fixed[c]})
        points = sorted({p for c in adjust for p in meas_of_kf[c]})
        pset = set(points)
        fixed_set = [c for c in range(C) if c not in adjust and any(p in pset for p in meas_of_kf[c])]
    b = Bundle(lib, g["width"], g["height"], max_iterations=max_iterations)
    cam_id, pt_id = {}, {}
    for c in adjust:
        cam_id[c] = b.AddCamera(g["cam_se3"][c], fixed[c])
    for c in fixed_set:
        cam_id[c] = b.AddCamera(g["cam_se3"][c], True)
    for p in points:
        pt_id[p] = b.AddPoint(g["points"][p])
    for c in range(C):
        if c not in cam_id:
            continue
        for p in sorted(meas_of_kf[c]):
            if p in pt_id:
                i = meas_of_kf[c][p]
                s = float(1 << int(g["meas_level"][i]))
                b.AddMeas(cam_id[c], pt_id[p], g["meas_uv"][i], s * s)
    acc = b.Compute()
    out_pts, out_cams = np.array(g["points"], np.float64).copy(), np.array(g["cam_se3"], np.float64).copy()
    if acc > 0:
        for p, i in pt_id.items():
            out_pts[p] = b.GetPoint(i)
        for c, i in cam_id.items():
            out_cams[c] = b.GetCamera(i)
        if mode == 1:
            flags["recent"] = False
        flags["full"] = False
    if b.Converged():
        flags["recent"] = True
        if mode == 0:
            flags["full"] = True
    view_of = {i: c for c, i in cam_id.items()}
    point_of = {i: p for p, i in pt_id.items()}
    bad = np.zeros(P, np.int32)
    queue, never = [], []
    for pi, ci in b.GetOutlierMeasurements():
        p, c = point_of[int(pi)], view_of[int(ci)]
        i = meas_of_kf[c][p]
        if len(kfs_of_pt[p]) <= 2 or g["meas_src"][i] == SRC_ROOT:
            bad[p] = 1
        else:
            if g["meas_src"][i] in (SRC_TRACKER, SRC_EPIPOLAR):
                queue.append((c, p))
            else:
                never.append((c, p))
            del meas_of_kf[c][p]
            kfs_of_pt[p].discard(c)
    b.close()
    nmeas = np.array([len(meas_of_kf[c]) for c in range(C)], np.int32)
    never = sorted(never, key=lambda cp: (cp[1], cp[0]))  # mapmaker_check lists them point by point, keyframes ascending
    return dict(points=out_pts, cams=out_cams, bad=bad, nmeas=nmeas, queue=np.array(queue, np.int32).reshape(-1, 2),
                never=np.array(never, np.int32).reshape(-1, 2),
                flags=np.array([flags["full"], flags["recent"], 0, 0], np.int32), accepted=acc)


def compare(got, exp, tol):
    np.testing.assert_allclose(got["points"], exp["points"], rtol=0, atol=tol)
    np.testing.assert_allclose(got["cams"], exp["cams"], rtol=0, atol=tol)
    assert np.array_equal(got["bad"], exp["bad"])
    assert np.array_equal(got["nmeas"], exp["nmeas"])
    assert np.array_equal(got["queue"], exp["queue"]), "failure queue (order included)"
    assert np.array_equal(got["never"], exp["never"])
    assert np.array_equal(got["flags"], exp["flags"])


# ---- MapMaker::AddPointsEpipolar (host mirror) against the reference's own AddPointEpipolar ---------------------
def write_epipolar_case(d, W, H, frames, poses, i_src, i_tgt, depth=(1.0, 0.3), wiggle=0.1):
    np.array([W, H], np.int32).tofile(d / "epi_dims.i32")
    np.ascontiguousarray(frames[i_src], np.uint8).tofile(d / "epi_src.u8")
    np.ascontiguousarray(frames[i_tgt], np.uint8).tofile(d / "epi_tgt.u8")
    np.ascontiguousarray(poses[i_src], np.float64).tofile(d / "epi_src_pose.f64")
    np.ascontiguousarray(poses[i_tgt], np.float64).tofile(d / "epi_tgt_pose.f64")
    np.array([depth[0], depth[1], wiggle], np.float64).tofile(d / "epi_depth.f64")


def read_epipolar_out(d):
    return dict(counts=np.fromfile(d / "epi_out_counts.i32", np.int32), ncand=np.fromfile(d / "epi_out_ncand.i32", np.int32),
                points=np.fromfile(d / "epi_out_points.f64").reshape(-1, 9), meas=np.fromfile(d / "epi_out_meas.f64").reshape(-1, 4),
                levels=np.fromfile(d / "epi_out_levels.i32", np.int32))


def epipolar_expected(search_lib, W, H, frames, poses, i_src, i_tgt, depth=(1.0, 0.3), wiggle=0.1, ref=None):
    """Candidates, accepted set and sub-pixel positions through `search_lib`'s C ABI; when `ref` (oracle/_ref) is
    given, also the world positions / pixel vectors the reference's own Triangulate + RefreshPixelVectors made."""
    import ctypes as C
    from ptam_cg_b200.capi import Tracker
    out = dict(counts=[], ncand=[], meas=[], levels=[], points=[])
    t = Tracker(search_lib, W, H, 1)
    kf = t.add_keyframe(frames[i_src])
    t.make_keyframes([frames[i_src]])
    cands = [r[1] for r in t.keyframe_rest(0, 70.0)]
    t.make_keyframes([frames[i_tgt]])
    if ref is not None:
        tr = Tracker(ref, W, H, 1)
        kfr = tr.add_keyframe(frames[i_src])
        tr.make_keyframes([frames[i_tgt]])
        last = ref.cdll.ref_tracker_epipolar_last_points
        last.restype = C.c_int
    for l in range(4):
        found, best, sub = t.epipolar_search(0, l, kf, poses[i_src], depth[0], depth[1], poses[i_tgt], wiggle, cands[l])
        out["ncand"].append(len(cands[l]))
        out["counts"].append(int(found.sum()))
        scale = 1 << l
        for i in np.flatnonzero(found):
            root = (np.asarray(cands[l][i], np.float64) + 0.5) * scale - 0.5   # Level::LevelZeroPos
            out["meas"].append([root[0], root[1], sub[i][0], sub[i][1]])
            out["levels"].append(l)
        if ref is not None:
            fr, _, subr = tr.epipolar_search(0, l, kfr, poses[i_src], depth[0], depth[1], poses[i_tgt], wiggle, cands[l])
            assert np.array_equal(fr, found)
            n = last(None, 0)
            buf = np.zeros((max(n, 1), 9))
            last(buf.ctypes.data_as(C.POINTER(C.c_double)), n)
            assert n == int(found.sum())
            out["points"].extend(buf[:n].tolist())
    out = {k: np.array(v) for k, v in out.items()}
    out["meas"] = out["meas"].reshape(-1, 4)
    out["points"] = out["points"].reshape(-1, 9)
    return out
