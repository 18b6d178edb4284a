"""The C++ host mirror's Bundle / KeyFrame / Tracker classes checked without a GPU: host_check.cc compiled
against the CPU oracle's identical ABI (tests/orc_alias.h) and compared with the ctypes binding of the same
library (tests/host_util.py; the GPU run against the CUDA library is tests/test_host_cpp_gpu.py)."""
import subprocess
from pathlib import Path

from host_util import check_host_classes
from oracle.binding import oracle_lib

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "ptam_cg_b200" / "host"


def test_host_classes_against_the_oracle_abi(tmp_path):
    out = tmp_path / "host_check_orc"
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", str(HOST), "-include", str(ROOT / "tests" / "orc_alias.h"),
           str(HOST / "host_check.cc"), "-o", str(out), "-L", str(ROOT / "oracle"), "-loracle", "-Wl,-rpath," + str(ROOT / "oracle")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    check_host_classes(oracle_lib(), out, tmp_path)
