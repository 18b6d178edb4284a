"""Pins the bundle-adjustment oracle against the reference's OWN code: oracle/_ref/libref_ptam.so is
src/Bundle.cc + src/ATANCamera.cc of /root/reference compiled in place (oracle/Makefile.ref) against
header stand-ins for TooN / libCVD / GVars3 (oracle/shim/).  With the platform atan on both sides the
oracle must reproduce the reference BIT FOR BIT (same accept/reject sequence, lambda trials, outlier
list in erase order, sigma^2, lambda, every camera and point); with the oracle's specified atan
(the numeric contract shared with the CUDA product) the same decisions and states within 1e-9.
What stays restated is only TooN's own arithmetic (SE3::exp, LDL^T Cholesky), see DESIGN.md §2."""
import numpy as np
import pytest

from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Bundle
from oracle.binding import oracle_lib, ref_lib

REF = ref_lib()
pytestmark = pytest.mark.skipif(REF is None, reason="oracle/_ref not built and /root/reference absent")


def _run(lib, g, **params):
    b = Bundle(lib, g["width"], g["height"], **params)
    b.add_graph(g)
    acc = b.Compute()
    s = b.stats()
    out = dict(acc=acc, trials=s.lambda_trials, converged=s.converged, hit_max=s.hit_max_iterations, sigma=s.sigma_squared,
               lam=s.lambda_, points=b.get_points(), cams=b.get_cameras(), outliers=b.GetOutlierMeasurements())
    b.close()
    return out


GRAPHS = [dict(n_cams=8, n_points=300, n_meas=1200, seed=1), dict(n_cams=20, n_points=1000, n_meas=5000, seed=2),
          dict(n_cams=50, n_points=5000, n_meas=20000, seed=42)]  # the last one is BASELINE config C3


@pytest.mark.parametrize("cfg", GRAPHS, ids=lambda c: f"{c['n_cams']}x{c['n_points']}x{c['n_meas']}")
def test_oracle_bit_identical_to_reference_bundle(cfg):
    g = synth.make_ba_graph(**cfg)
    o, r = _run(oracle_lib(libm_atan=True), g), _run(REF, g)
    for k in ("acc", "trials", "converged", "hit_max", "sigma", "lam"):
        assert o[k] == r[k], k
    assert np.array_equal(o["outliers"], r["outliers"])
    assert np.array_equal(o["points"], r["points"]) and np.array_equal(o["cams"], r["cams"])
    assert o["acc"] >= 5 and len(o["outliers"]) > 0


@pytest.mark.parametrize("cfg", GRAPHS[:2], ids=lambda c: f"{c['n_cams']}x{c['n_points']}x{c['n_meas']}")
def test_spec_atan_oracle_within_tolerance_of_reference(cfg):
    g = synth.make_ba_graph(**cfg)
    o, r = _run(oracle_lib(), g), _run(REF, g)
    assert (o["acc"], o["trials"]) == (r["acc"], r["trials"])
    assert np.array_equal(o["outliers"], r["outliers"])
    assert np.allclose(o["points"], r["points"], rtol=0, atol=1e-8) and np.allclose(o["cams"], r["cams"], rtol=0, atol=1e-8)


@pytest.mark.parametrize("est", [1, 2], ids=["Cauchy", "Huber"])
def test_other_mestimators_match_reference(est):
    g = synth.make_ba_graph(8, 300, 1200, seed=3)
    o, r = _run(oracle_lib(libm_atan=True), g, mestimator=est), _run(REF, g, mestimator=est)
    assert (o["acc"], o["trials"], o["lam"]) == (r["acc"], r["trials"], r["lam"])
    assert np.array_equal(o["outliers"], r["outliers"])
    assert np.array_equal(o["points"], r["points"]) and np.array_equal(o["cams"], r["cams"])


@pytest.mark.parametrize("max_it", [1, 2, 3, 5])
def test_truncated_runs_match_reference(max_it):
    """The reference cannot be stepped from outside; capping Bundle.MaxIterations exposes the state after
    the first lambda trials instead."""
    g = synth.make_ba_graph(12, 500, 2500, seed=4)
    o, r = _run(oracle_lib(libm_atan=True), g, max_iterations=max_it), _run(REF, g, max_iterations=max_it)
    assert o["trials"] == r["trials"] == max_it and o["acc"] == r["acc"]
    assert np.array_equal(o["points"], r["points"]) and np.array_equal(o["cams"], r["cams"])


def test_edge_cases_match_reference():
    g = synth.make_ba_graph(6, 120, 500, seed=5)
    g["cam_fixed"] = np.asarray(g["cam_fixed"]).copy()
    g["cam_fixed"][[0, 2]] = 1                      # two fixed cameras
    g["points"] = np.asarray(g["points"]).copy()
    g["points"][7] = np.nan                         # NaN point is zeroed (Bundle.cc:70-74)
    g["points"][11] = g["cam_se3"][1][9:12] * -5.0  # a point far off: behind some cameras
    o, r = _run(oracle_lib(libm_atan=True), g), _run(REF, g)
    assert (o["acc"], o["trials"]) == (r["acc"], r["trials"])
    assert np.array_equal(o["outliers"], r["outliers"])
    assert np.array_equal(o["points"], r["points"], equal_nan=True) and np.array_equal(o["cams"], r["cams"], equal_nan=True)


def test_recompute_on_the_same_object_matches_reference():
    """Persistent graph: Compute, then Compute again on the state it left (adjusted
    poses / points, outliers erased) — the reference object itself can be computed twice (built with
    NDEBUG like a release build: GenerateOffDiagScripts asserts on the erased measurements otherwise)."""
    g = synth.make_ba_graph(12, 500, 2500, seed=6)
    out = []
    for lib in (oracle_lib(libm_atan=True), REF):
        b = Bundle(lib, g["width"], g["height"], max_iterations=4)
        b.add_graph(g)
        a1 = b.Compute()
        o1 = b.GetOutlierMeasurements()
        b.update_camera(3, b.GetCamera(3))            # a no-op update through the new entry points
        p7 = b.GetPoint(7) + 1e-3
        b.update_point(7, p7)
        a2 = b.Recompute()
        s = b.stats()
        out.append((a1, o1, a2, s.lambda_trials, b.GetOutlierMeasurements(), b.get_points(), b.get_cameras()))
        b.close()
    o, r = out
    assert o[0] == r[0] and np.array_equal(o[1], r[1]) and len(o[1]) > 0
    assert (o[2], o[3]) == (r[2], r[3]) and np.array_equal(o[4], r[4])
    assert np.array_equal(o[5], r[5]) and np.array_equal(o[6], r[6])


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15, 16])
def test_seed_sweep_bit_identical_to_reference(seed):
    """Different shapes, outlier fractions and M-estimators per seed: the oracle must follow the reference's own
    Bundle.cc bit for bit on every one of them (decisions, outlier order, every state)."""
    rng = np.random.default_rng(seed)
    n_cams = int(rng.integers(4, 25))
    n_points = int(rng.integers(80, 900))
    n_meas = int(n_points * rng.uniform(2.2, 4.5))
    g = synth.make_ba_graph(n_cams, n_points, n_meas, seed=seed, outlier_frac=float(rng.choice([0.0, 0.02, 0.1])))
    prm = dict(mestimator=int(rng.integers(0, 3)), max_iterations=int(rng.integers(3, 21)))
    o, r = _run(oracle_lib(libm_atan=True), g, **prm), _run(REF, g, **prm)
    for k in ("acc", "trials", "converged", "hit_max", "sigma", "lam"):
        assert o[k] == r[k], k
    assert np.array_equal(o["outliers"], r["outliers"])
    assert np.array_equal(o["points"], r["points"]) and np.array_equal(o["cams"], r["cams"])
