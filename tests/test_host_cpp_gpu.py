"""The C++ host mirror (ptam_cg_b200/host: Bundle / KeyFrame / Tracker with TooN/CVD-style types)
run end to end on the GPU through host_check, compared with the same inputs driven through the
ctypes binding (which the other GPU tests hold bit-exact / within tolerance against the oracle)."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "ptam_cg_b200" / "host"


def test_host_mirror_builds():
    r = subprocess.run(["make", "-C", str(HOST)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (HOST / "host_check").exists()


@pytest.mark.gpu
def test_host_classes_match_capi(product, tmp_path):
    subprocess.run(["make", "-C", str(HOST)], check=True, capture_output=True)
    from host_util import check_host_classes
    check_host_classes(product, HOST / "host_check", tmp_path)
