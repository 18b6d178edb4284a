"""CPU-only: libptam_b200.so loads and exports every symbol include/ptam_b200.h declares, the
oracle exports the mirrored orc_* set, and creating a handle without a GPU fails loudly."""
import re
from pathlib import Path

import numpy as np
import pytest

from ptam_cg_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "ptam_b200.h").read_text()
    return sorted(set(re.findall(r"\b(ptam_(?:tracker|bundle|global|nccl|patch|pose)_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(product):
    names = declared_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(product.cdll, n)]
    assert not missing, missing


def test_binding_lists_match_header():
    hdr = {n[len("ptam_"):] for n in declared_symbols()}
    listed = set(capi.TRACKER_SYMBOLS + capi.BUNDLE_SYMBOLS + capi.PRODUCT_ONLY_SYMBOLS)
    assert hdr == listed, hdr ^ listed


def test_oracle_mirrors_the_abi(oracle):
    for n in capi.TRACKER_SYMBOLS + capi.BUNDLE_SYMBOLS:
        assert oracle.has(n), n


def test_struct_layouts_match_header_sizes():
    import ctypes
    assert ctypes.sizeof(capi.TrackerParams) == 8 * 4 + 3 * 8 + 2 * 4 + 8
    assert ctypes.sizeof(capi.TrackerState) == 21 * 8 + 4 * 4
    assert ctypes.sizeof(capi.TrackResult) == 14 * 8 + 26 * 4 + 8  # 23 ints + recovery, reloc_keyframe, reserved1; reloc_score
    assert ctypes.sizeof(capi.BundleParams) == 2 * 4 + 2 * 8
    assert ctypes.sizeof(capi.BundleStats) == 6 * 4 + 4 * 8


def test_no_cpu_fallback(product):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.PtamError, match="no CUDA device|CPU fallback"):
        capi.Tracker(product, 640, 480)
    with pytest.raises(capi.PtamError):
        capi.Bundle(product, 640, 480)


def test_product_never_imports_oracle():
    for f in list((ROOT / "ptam_cg_b200").rglob("*.py")) + list((ROOT / "ptam_cg_b200").rglob("*.cu")) + \
            list((ROOT / "ptam_cg_b200").rglob("*.cuh")) + list((ROOT / "ptam_cg_b200").rglob("*.h")):
        text = f.read_text()
        assert "oracle.binding" not in text and "liboracle" not in text and "oracle_math" not in text, f
