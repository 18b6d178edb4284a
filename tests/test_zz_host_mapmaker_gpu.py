"""The MapMaker host mirror (ptam_cg_b200/host/MapMaker.h) on the GPU: mapmaker_check linked against the CUDA
library, its map compared with the reference's control flow driven over the same C ABI from Python
(tests/mapmaker_util.py; the CPU test of the same source against the oracle ABI is test_host_mapmaker_cpu.py).
Two runs of the CUDA library differ by the order of its atomic sums, hence the 1e-7 on the states (as in
tests/test_host_cpp_gpu.py); the discrete bookkeeping must be equal.  Sorted last: new coverage of the callers'
side of path B, not part of the core suite's `-x` chain."""
import subprocess
from pathlib import Path

import pytest

import mapmaker_util as mu

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "ptam_cg_b200" / "host"

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1], ids=["BundleAdjustAll", "BundleAdjustRecent"])
def test_mapmaker_mirror_on_the_device(product, tmp_path, mode):
    r = subprocess.run(["make", "-C", str(HOST)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    g = mu.make_map()
    mu.write_map(g, tmp_path, mode, 20)
    r = subprocess.run([str(HOST / "mapmaker_check"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = mu.read_map(tmp_path, len(g["cam_fixed"]), len(g["points"]))
    exp = mu.expected(product, g, mode, 20)
    assert exp["accepted"] > 0
    mu.compare(got, exp, tol=1e-7)


def test_add_points_epipolar_on_the_device(product, oracle, tmp_path):
    """MapMaker::AddPointsEpipolar of the host mirror with the CUDA library behind it, against the same source
    compiled over the CPU oracle's ABI (which tests/test_host_mapmaker_cpu.py holds against the reference's own
    MapMaker::AddPointEpipolar): same accepted candidates, measurements to 1e-9 px, new points to 1e-6."""
    import numpy as np
    from ptam_cg_b200 import synth
    r = subprocess.run(["make", "-C", str(HOST)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    orc_bin = tmp_path / "mapmaker_check_orc"
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", str(HOST), "-include", str(ROOT / "tests" / "orc_alias.h"),
           str(HOST / "mapmaker_check.cc"), "-o", str(orc_bin), "-L", str(ROOT / "oracle"), "-loracle", "-Wl,-rpath," + str(ROOT / "oracle")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    W, H = 320, 240
    frames, poses = synth.render_sequence(W, H, 40)
    outs = []
    for binary, sub in ((HOST / "mapmaker_check", "gpu"), (orc_bin, "cpu")):
        d = tmp_path / sub
        d.mkdir()
        mu.write_epipolar_case(d, W, H, frames, poses, 0, 30)
        r = subprocess.run([str(binary), str(d), "epi"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(mu.read_epipolar_out(d))
    g, c = outs
    assert np.array_equal(g["ncand"], c["ncand"]) and np.array_equal(g["counts"], c["counts"]) and g["counts"].sum() > 100
    assert np.array_equal(g["levels"], c["levels"])
    assert np.array_equal(g["meas"][:, :2], c["meas"][:, :2])
    np.testing.assert_allclose(g["meas"][:, 2:], c["meas"][:, 2:], rtol=0, atol=1e-9)
    np.testing.assert_allclose(g["points"], c["points"], rtol=0, atol=1e-6)


def test_refind_in_single_keyframe_on_the_device(product, tmp_path):
    """MapMaker::ReFindInSingleKeyFrame of the host mirror with the CUDA library behind it, against the same C ABI
    driven from Python + the reference's bookkeeping (the CPU twin is in tests/test_host_mapmaker_cpu.py)."""
    import numpy as np
    from test_host_mapmaker_cpu import _refind_case
    r = subprocess.run(["make", "-C", str(HOST)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    n, exp, pos, n_new = _refind_case(tmp_path, product)
    r = subprocess.run([str(HOST / "mapmaker_check"), str(tmp_path), "refind"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "rf_out_points.i32", np.int32).reshape(n, 5)
    gpos = np.fromfile(tmp_path / "rf_out_pos.f64").reshape(n, 2)
    counts = np.fromfile(tmp_path / "rf_out_counts.i32", np.int32)
    assert list(counts[:2]) == [n_new, 0] and n_new > 100
    assert np.array_equal(got, exp)
    np.testing.assert_allclose(gpos, pos, rtol=0, atol=1e-9)
    n_queued, n_second, n_same, left = counts[2:]   # ReFindFromFailureQueue round trip
    assert n_queued > 10 and n_second == n_queued and n_same == n_queued and left == 0


def test_add_keyframe_from_top_of_queue_on_the_device(product, tmp_path):
    """MapMaker::AddKeyFrame + AddKeyFrameFromTopOfQueue of the host mirror (re-find, ThinCandidates, epipolar search
    on levels 3, 0, 1, 2 in the nearest keyframe) with the CUDA library behind it; expectation from the same C ABI
    driven from Python (CPU twin: tests/test_host_mapmaker_cpu.py::test_add_keyframe_from_top_of_queue)."""
    from test_host_mapmaker_cpu import _add_keyframe_case, _check_add_keyframe
    r = subprocess.run(["make", "-C", str(HOST)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    case = _add_keyframe_case(tmp_path, product)
    r = subprocess.run([str(HOST / "mapmaker_check"), str(tmp_path), "addkf"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    _check_add_keyframe(tmp_path, *case, tol=1e-9)


def test_tracker_and_mapmaker_loop_on_the_device(product, tmp_path):
    """Both host mirrors in the reference's loop with the CUDA library behind them (track, hand a keyframe over, new
    points by epipolar search, track on, re-find, bundle-adjust): same invariants as the CPU twin."""
    from test_host_mapmaker_cpu import _loop_case, _check_loop
    r = subprocess.run(["make", "-C", str(HOST)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    case = _loop_case(tmp_path)
    r = subprocess.run([str(HOST / "mapmaker_check"), str(tmp_path), "loop"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    _check_loop(tmp_path, *case)


def test_tracker_consults_the_map_maker_on_the_device(product, tmp_path):
    """The Tracker mirror with a map maker attached and the CUDA library behind it: tracking quality, lost-frame
    counter and keyframe queue after every frame, against the reference's OWN Tracker + MapMaker (oracle/_ref), with
    the default thresholds (discrete outcomes far from their thresholds on this sequence)."""
    import numpy as np
    from oracle.binding import ref_lib
    from test_host_mapmaker_cpu import _heuristics_case, _check_heuristics_against_reference
    ref = ref_lib()
    if ref is None or not hasattr(ref.cdll, "ref_tracker_mapmaker_ctl"):
        pytest.skip("oracle/_ref not present")
    r = subprocess.run(["make", "-C", str(HOST)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    params = (0.3, 0.13, 0.1)
    case = _heuristics_case(tmp_path, params)
    nfr = case[3]
    r = subprocess.run([str(HOST / "mapmaker_check"), str(tmp_path), "heur"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "hq_out.i32", np.int32).reshape(nfr, 4)
    _check_heuristics_against_reference(ref, params, *case, got)
