"""More parity cases for both paths on inputs none of the other GPU tests use (other scene seeds = the C5
stream seeds of SURVEY 8d, other graph shapes / outlier fractions / M-estimators).  Sorted last on purpose:
these widen the coverage of the product against the oracle and must not mask the core suite under `-x`.
Same bars as tests/test_tracker_gpu.py and tests/test_bundle_gpu.py."""
import numpy as np
import pytest

from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Bundle, Tracker
from oracle.binding import detect_with

from test_tracker_gpu import _compare_frame, _compare_levels

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [20260102, 20260105])
def test_track_frames_other_scenes_from_identical_state(oracle, product, seed):
    W, H = 320, 240
    frames, poses = synth.render_sequence(W, H, 12, seed=seed)
    cam = synth.AtanCamera(W, H)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle, W, H), cam, kf_indices=(0, 6), per_level=(150, 80, 40, 20))
    pair = []
    for lib in (oracle, product):
        t = Tracker(lib, W, H, 1, use_rotation_estimator=0, mestimator=seed % 3)
        for k in kfs:
            t.add_keyframe(k)
        t.set_map(0, m)
        pair.append(t)
    o, p = pair
    start = synth.perturb_pose(poses[1], np.random.default_rng(seed))
    o.set_state(0, pose12=start, msd=0.02); p.set_state(0, pose12=start, msd=0.02)
    found = 0
    for f in range(1, 11):
        p.set_state(0, state=o.get_state(0))  # per-step parity from identical state
        ro = o.track_frames([frames[f]])[0]
        rp = p.track_frames([frames[f]])[0]
        _compare_levels(o, p)
        _compare_frame(o, p, ro, rp)
        found += sum(ro.meas_found)
    assert found > 500


@pytest.mark.parametrize("seed", [21, 22, 23, 24])
def test_first_lm_step_seed_sweep(oracle, product, seed):
    rng = np.random.default_rng(seed)
    n_cams = int(rng.integers(4, 40))
    n_points = int(rng.integers(100, 3000))
    n_meas = int(n_points * rng.uniform(2.2, 4.5))
    g = synth.make_ba_graph(n_cams, n_points, n_meas, seed=seed, outlier_frac=float(rng.choice([0.0, 0.02, 0.1])))
    prm = dict(mestimator=int(rng.integers(0, 3)))
    o, p = Bundle(oracle, g["width"], g["height"], **prm), Bundle(product, g["width"], g["height"], **prm)
    o.add_graph(g); p.add_graph(g)
    o.begin(); p.begin()
    o.lm_step(); p.lm_step()   # both start from bit-identical state: the tight comparison
    so, sp = o.stats(), p.stats()
    for f in ("accepted", "lambda_trials", "lm_steps", "converged", "hit_max_iterations", "n_outliers"):
        assert getattr(so, f) == getattr(sp, f), f
    for f in ("sigma_squared", "lambda_", "last_error", "last_new_error"):
        np.testing.assert_allclose(getattr(sp, f), getattr(so, f), rtol=1e-10, err_msg=f)
    n = 6 * int((g["cam_fixed"] == 0).sum())
    So, eo = o.reduced_system(n)
    Sp, ep = p.reduced_system(n)
    np.testing.assert_allclose(Sp, So, atol=1e-12 * np.abs(So).max(), rtol=0)
    np.testing.assert_allclose(ep, eo, atol=1e-12 * np.abs(eo).max(), rtol=0)
    assert np.array_equal(o.GetOutlierMeasurements(), p.GetOutlierMeasurements())
    np.testing.assert_allclose(p.get_points(), o.get_points(), atol=1e-9, rtol=0)
    np.testing.assert_allclose(p.get_cameras(), o.get_cameras(), atol=1e-9, rtol=0)


@pytest.mark.parametrize("n_cams", [33, 65], ids=["n192", "n384"])
def test_reduced_system_of_whole_panels(oracle, product, n_cams):
    """6 x (n_cams - 1) free-camera rows = a multiple of the 64-wide LDL^T panel: the last panel is a full one with
    nothing below it (bulk-copied operands, no rows to solve) — a shape no other graph in the suite has."""
    g = synth.make_ba_graph(n_cams, 1500, 6000, seed=50 + n_cams)
    n = 6 * int((g["cam_fixed"] == 0).sum())
    assert n % 64 == 0
    o, p = Bundle(oracle, g["width"], g["height"]), Bundle(product, g["width"], g["height"])
    o.add_graph(g); p.add_graph(g)
    o.begin(); p.begin()
    o.lm_step(); p.lm_step()
    so, sp = o.stats(), p.stats()
    assert (so.accepted, so.lambda_trials, so.n_outliers) == (sp.accepted, sp.lambda_trials, sp.n_outliers)
    np.testing.assert_allclose(sp.last_new_error, so.last_new_error, rtol=1e-10)
    assert np.array_equal(o.GetOutlierMeasurements(), p.GetOutlierMeasurements())
    np.testing.assert_allclose(p.get_points(), o.get_points(), atol=1e-9, rtol=0)
    np.testing.assert_allclose(p.get_cameras(), o.get_cameras(), atol=1e-9, rtol=0)
