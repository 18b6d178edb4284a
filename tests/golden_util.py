"""Shared by the CPU (oracle) and GPU (product) golden-fixture tests."""
from pathlib import Path

import numpy as np

from ptam_cg_b200.capi import Bundle, Tracker

GOLD = Path(__file__).resolve().parent / "golden"


def check_tracker_golden(lib, pose_tol, subpix_tol):
    z = np.load(GOLD / "tracker_192x144.npz")
    frames = z["frames"]
    W, H = frames.shape[2], frames.shape[1]
    m = {k[4:]: z[k] for k in z.files if k.startswith("map_")}
    t = Tracker(lib, W, H, 1)
    for i in z["kf_indices"]:
        t.add_keyframe(frames[int(i)])
    t.set_map(0, m)
    t.set_state(0, pose12=z["start"], msd=0.02)
    r = t.track_frames([frames[4]])[0]
    for l in range(4):
        _, xy, lut = t.get_level(0, l)
        assert np.array_equal(xy, z[f"corners{l}"]) and np.array_equal(lut, z[f"lut{l}"])
    p = t.get_points(0)
    assert np.array_equal(p["flags"], z["flags"]) and np.array_equal(p["level"], z["level"])
    assert np.array_equal(t.get_iteration_set(0), z["iteration_set"])
    assert np.array_equal(t.get_templates(0)[0], z["templates"])
    assert list(r.meas_attempted) == list(z["attempted"]) and list(r.meas_found) == list(z["found"])
    assert r.did_coarse == int(z["did_coarse"]) and [r.n_coarse, r.n_level3, r.n_fine] == list(z["n_sets"])
    sub = (z["flags"] & 16) != 0
    found = (z["flags"] & 8) != 0
    assert np.array_equal(p["v2_found"][found & ~sub], z["v2_found"][found & ~sub])
    np.testing.assert_allclose(p["v2_found"][sub], z["v2_found"][sub], atol=subpix_tol, rtol=0)
    np.testing.assert_allclose(np.array(r.se3_cam_from_world), z["pose"], atol=pose_tol, rtol=0)
    assert sum(r.meas_found) > 40


def check_bundle_golden(lib, tol):
    z = np.load(GOLD / "bundle_6x120x480.npz")
    g = {k[2:]: z[k] for k in z.files if k.startswith("g_")}
    b = Bundle(lib, int(g["width"]), int(g["height"]))
    b.add_graph(g)
    assert b.Compute() == int(z["accepted"])
    s = b.stats()
    assert s.lambda_trials == int(z["lambda_trials"]) and s.n_outliers == int(z["n_outliers"])
    assert np.array_equal(b.GetOutlierMeasurements(), z["outliers"])
    np.testing.assert_allclose(s.sigma_squared, float(z["sigma_squared"]), rtol=max(tol, 1e-12) * 1e3)
    np.testing.assert_allclose(b.get_points(), z["points"], atol=tol, rtol=0)
    np.testing.assert_allclose(b.get_cameras(), z["cameras"], atol=tol, rtol=0)
