"""The CUDA path reproduces the committed golden fixtures (integers exactly, f64 to tolerance)."""
import pytest

from golden_util import check_bundle_golden, check_tracker_golden

pytestmark = pytest.mark.gpu


def test_product_tracker_golden(product):
    check_tracker_golden(product, pose_tol=1e-9, subpix_tol=1e-6)


def test_product_bundle_golden(product):
    check_bundle_golden(product, tol=1e-8)
