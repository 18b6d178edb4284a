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
