"""The oracle reproduces the committed golden fixtures bit-for-bit (regression pin, CPU only)."""
from golden_util import check_bundle_golden, check_tracker_golden


def test_oracle_tracker_golden(oracle):
    check_tracker_golden(oracle, pose_tol=0.0, subpix_tol=0.0)


def test_oracle_bundle_golden(oracle):
    check_bundle_golden(oracle, tol=0.0)
