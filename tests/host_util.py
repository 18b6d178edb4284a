"""Shared by the GPU and CPU tests of the C++ host mirror (ptam_cg_b200/host: Bundle / KeyFrame / Tracker with
TooN/CVD-style types): runs host_check on raw arrays and compares with the same inputs driven through the ctypes
binding of the SAME library (the CUDA product on the GPU; on CPU the oracle, whose identical ABI the host sources
are compiled against through tests/orc_alias.h)."""
import subprocess

import numpy as np

from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Bundle, Tracker


def check_host_classes(lib, binary, tmp_path):
    product = lib
    # ---- inputs
    g = synth.make_ba_graph(6, 150, 600, seed=5)
    np.ascontiguousarray(g["cam_se3"], np.float64).tofile(tmp_path / "ba_cams.f64")
    np.ascontiguousarray(g["cam_fixed"], np.int32).tofile(tmp_path / "ba_fixed.i32")
    np.ascontiguousarray(g["points"], np.float64).tofile(tmp_path / "ba_pts.f64")
    np.ascontiguousarray(g["meas_cam"], np.int32).tofile(tmp_path / "ba_mcam.i32")
    np.ascontiguousarray(g["meas_point"], np.int32).tofile(tmp_path / "ba_mpt.i32")
    np.ascontiguousarray(g["meas_uv"], np.float64).tofile(tmp_path / "ba_uv.f64")
    np.ascontiguousarray(g["meas_sigma_sq"], np.float64).tofile(tmp_path / "ba_s2.f64")
    assert (g["width"], g["height"]) == (640, 480)
    W, H, NF = 320, 240, 4
    frames, poses = synth.render_sequence(W, H, 8)
    cam = synth.AtanCamera(W, H)
    det = Tracker(product, W, H, 1)

    def detect(image):
        det.make_keyframes([image])
        return [det.get_level(0, l)[:2] for l in range(4)]

    kfs, m = synth.build_map(frames, poses, detect, cam, kf_indices=(0, 4), per_level=(150, 80, 40, 20))
    pose0 = synth.perturb_pose(poses[1], np.random.default_rng(3))
    np.array([W, H, len(kfs), len(m["src_kf"]), NF], np.int32).tofile(tmp_path / "trk_dims.i32")
    np.ascontiguousarray(np.stack(kfs), np.uint8).tofile(tmp_path / "trk_kf.u8")
    np.ascontiguousarray(frames[1:1 + NF], np.uint8).tofile(tmp_path / "trk_frames.u8")
    for name, key, dt in (("trk_world.f64", "world_pos", np.float64), ("trk_right.f64", "pixel_right_w", np.float64),
                          ("trk_down.f64", "pixel_down_w", np.float64), ("trk_srckf.i32", "src_kf", np.int32),
                          ("trk_srclevel.i32", "src_level", np.int32), ("trk_center.i32", "ir_center", np.int32)):
        np.ascontiguousarray(m[key], dt).tofile(tmp_path / name)
    np.ascontiguousarray(pose0, np.float64).tofile(tmp_path / "trk_pose0.f64")
    pf_pose = synth.perturb_pose(poses[NF], np.random.default_rng(5))   # for class PatchFinder: the last frame is frames[NF]
    PF_N, PF_RANGE, PF_ITS = 60, 10, 8
    np.ascontiguousarray(pf_pose, np.float64).tofile(tmp_path / "pf_pose.f64")
    np.array([PF_N, PF_RANGE, PF_ITS], np.int32).tofile(tmp_path / "pf_cfg.i32")
    # ---- C++ classes
    r = subprocess.run([str(binary), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # ---- same inputs through the ctypes binding
    b = Bundle(product, 640, 480)
    b.add_graph(g)
    acc = b.Compute()
    meta = np.fromfile(tmp_path / "ba_out_meta.i32", np.int32)
    assert meta[0] == acc and meta[1] == int(b.Converged()) and meta[2] == len(b.GetOutlierMeasurements())
    assert np.allclose(np.fromfile(tmp_path / "ba_out_pts.f64").reshape(-1, 3), b.get_points(), atol=1e-7)
    assert np.allclose(np.fromfile(tmp_path / "ba_out_cams.f64").reshape(-1, 12), b.get_cameras(), atol=1e-7)
    assert np.array_equal(np.fromfile(tmp_path / "ba_out_outliers.i32", np.int32).reshape(-1, 2), b.GetOutlierMeasurements())
    t = Tracker(product, W, H, 1)
    for k in kfs:
        t.add_keyframe(k)
    t.set_map(0, m)
    t.set_state(0, pose12=pose0, velocity=np.zeros(6), msd=0.0)
    out_poses = np.fromfile(tmp_path / "trk_out_poses.f64").reshape(NF, 12)
    out_found = np.fromfile(tmp_path / "trk_out_found.i32", np.int32).reshape(NF, 4)
    for f in range(NF):
        res = t.track_frames([frames[1 + f]])[0]
        assert np.allclose(out_poses[f], np.array(res.se3_cam_from_world), atol=1e-9)
        assert list(out_found[f]) == list(res.meas_found)
    last = np.fromfile(tmp_path / "trk_out_last.i32", np.int32)
    pts = t.get_points(0)
    assert last[0] == int(((pts["flags"] & 8) != 0).sum())
    assert list(last[1:]) == [len(t.get_level(0, l)[1]) for l in range(4)]
    # ---- class PatchFinder (host/PatchFinder.h), step by step, against the batched C-ABI entry on the same inputs
    u = Tracker(product, W, H, 1)
    for k in kfs:
        u.add_keyframe(k)
    u.set_map(0, m)
    u.make_keyframes([frames[NF]])
    u.patch_search_batch(np.asarray(pf_pose).reshape(1, 12), PF_RANGE, 0)
    coarse = u.patch_results(0)
    tmpl, sums = u.get_templates(0)
    u.patch_search_batch(np.asarray(pf_pose).reshape(1, 12), PF_RANGE, PF_ITS)
    fine = u.patch_results(0)
    oi = np.fromfile(tmp_path / "pf_out.i32", np.int32).reshape(-1, 7)
    od = np.fromfile(tmp_path / "pf_out.f64").reshape(-1, 6)
    ot = np.fromfile(tmp_path / "pf_out_tmpl.u8", np.uint8).reshape(-1, 64)
    n = len(oi)
    assert n == PF_N
    assert np.array_equal(oi[:, 0], coarse["level"][:n])
    assert np.array_equal(od[:, :4], coarse["warp_inverse"][:n])
    searched = (oi[:, 0] >= 0) & (oi[:, 1] == 0)
    assert np.array_equal(oi[:, 1][oi[:, 0] >= 0], coarse["template_bad"][:n][oi[:, 0] >= 0])
    assert np.array_equal(oi[:, 2][searched], coarse["found"][:n][searched])
    fc = searched & (oi[:, 2] == 1)
    assert fc.sum() > 10
    assert np.array_equal(oi[fc][:, 3:5], coarse["pos"][:n][fc].astype(np.int32))
    assert np.array_equal(ot[searched], tmpl[:n][searched]) and np.array_equal(oi[:, 5][searched], sums[:n, 0][searched])
    assert np.array_equal(oi[:, 6][fc], fine["found"][:n][fc])
    conv = fc & (oi[:, 6] == 1)
    assert np.allclose(od[conv][:, 4:6], fine["pos"][:n][conv], atol=1e-9)
    bi = np.fromfile(tmp_path / "pf_batch.i32", np.int32).reshape(-1, 2)
    bd = np.fromfile(tmp_path / "pf_batch.f64").reshape(-1, 2)
    assert np.array_equal(bi[:, 0], fine["level"]) and np.array_equal(bi[:, 1], fine["found"])
    assert np.allclose(bd, fine["pos"], atol=1e-9)
    det.make_keyframes([kfs[0]])
    assert np.array_equal(np.fromfile(tmp_path / "trk_out_kf0_corners.i32", np.int32).reshape(-1, 2), det.get_level(0, 0)[1])
    mx, cx, _ = det.keyframe_rest(0)[0]
    rest = np.fromfile(tmp_path / "trk_out_kf0_rest.i32", np.int32)
    assert rest[0] == len(mx) and rest[1] == len(cx) and np.array_equal(rest[2:].reshape(-1, 2), cx)
