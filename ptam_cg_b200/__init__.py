"""ptam_cg_b200 — B200-native PTAM hot paths (per-frame tracker + bundle adjuster).

Product = hand-written sm_100a CUDA behind the C-ABI of include/ptam_b200.h
(csrc/libptam_b200.so), the C++ host mirror of the reference classes (host/), and this thin
ctypes binding used by tests and bench.py.  No CPU fallback: the CUDA library must be built.
"""
from .capi import (Bundle, Lib, PtamError, Tracker, product_lib, CAMERA_PARAMS, LEVELS,  # noqa: F401
                   PT_FOUND, PT_IN_IMAGE, PT_IN_PVS, PT_SEARCHED, PT_SUBPIX, PT_TEMPLATE_BAD)
