import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built libraries (they are git-ignored): build them once, as
    # __graft_entry__.build() does (nvcc cross-compiles sm_100a without a GPU)
    lib = ROOT / "ptam_cg_b200" / "csrc" / "libptam_b200.so"
    host = ROOT / "ptam_cg_b200" / "host" / "host_check"
    orc = ROOT / "oracle" / "liboracle.so"
    if not (lib.exists() and host.exists() and orc.exists()):
        import __graft_entry__
        __graft_entry__.build()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import oracle_lib
    return oracle_lib()


@pytest.fixture(scope="session")
def product():
    from ptam_cg_b200.capi import product_lib
    return product_lib()


@pytest.fixture(scope="session")
def seq640():
    """48 synthetic 640x480 frames + ground-truth poses (SURVEY.md §8d scene, shorter run)."""
    from ptam_cg_b200 import synth
    frames, poses = synth.render_sequence(640, 480, 48)
    return frames, poses


@pytest.fixture(scope="session")
def map640(seq640, oracle):
    """~1000-point map built from 4 source keyframes with the oracle's FAST corners."""
    from ptam_cg_b200 import synth
    from ptam_cg_b200.capi import Tracker
    from oracle.binding import detect_with
    frames, poses = seq640
    cam = synth.AtanCamera(640, 480)
    kfs, m = synth.build_map(frames, poses, detect_with(Tracker, oracle, 640, 480), cam, kf_indices=(0, 12, 24, 36))
    return kfs, m
