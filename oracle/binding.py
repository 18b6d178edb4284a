"""TEST INFRASTRUCTURE — loads liboracle.so (the CPU restatement of the reference's hot paths)
through the same generic ctypes binding the product uses.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference leg may import this module.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from ptam_cg_b200.capi import Lib

ORACLE_DIR = Path(__file__).resolve().parent


def build(force=False):
    so = ORACLE_DIR / "liboracle.so"
    srcs = [ORACLE_DIR / n for n in ("oracle_tracker.cpp", "oracle_bundle.cpp", "oracle_math.h")]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return so


_lib = {}


def oracle_lib(libm_atan=False) -> Lib:
    name = "liboracle_libm.so" if libm_atan else "liboracle.so"
    if name not in _lib:
        build()
        lib = Lib(ORACLE_DIR / name, "orc_")
        c = lib.cdll
        d, i, P = C.c_double, C.c_int, C.POINTER
        c.orc_atan.restype, c.orc_atan.argtypes = d, [d]
        c.orc_se3_exp.restype, c.orc_se3_exp.argtypes = None, [P(d), P(d)]
        c.orc_se3_ln.restype, c.orc_se3_ln.argtypes = None, [P(d), P(d)]
        c.orc_cam_project.restype, c.orc_cam_project.argtypes = None, [P(d), i, i, P(d), P(d), P(d), P(i)]
        c.orc_cam_unproject.restype, c.orc_cam_unproject.argtypes = None, [P(d), i, i, P(d), P(d)]
        c.orc_cam_largest_radius.restype, c.orc_cam_largest_radius.argtypes = d, [P(d), i, i]
        c.orc_zmssd.restype, c.orc_zmssd.argtypes = i, [P(C.c_uint8), i, i, i, i, P(C.c_uint8)]
        c.orc_fast10_bruteforce.restype = i
        c.orc_fast10_bruteforce.argtypes = [P(C.c_uint8), i, i, i, P(C.c_int32), i]
        _lib[name] = lib
    return _lib[name]


def detect_with(tracker_cls, lib, width, height):
    """detect(image) callable for synth.build_map built on a (oracle or product) Tracker."""
    trk = tracker_cls(lib, width, height, 1)

    def detect(image):
        trk.make_keyframes([image])
        out = []
        for l in range(4):
            pix, xy, _ = trk.get_level(0, l)
            out.append((pix, xy))
        return out

    return detect


def fast10_bruteforce(im, t):
    lib = oracle_lib()
    im = np.ascontiguousarray(im, np.uint8)
    h, w = im.shape
    cap = max(w * h, 1)
    xy = np.zeros((cap, 2), np.int32)
    n = lib.cdll.orc_fast10_bruteforce(im.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, t,
                                       xy.ctypes.data_as(C.POINTER(C.c_int32)), cap)
    return xy[:n]


REFERENCE_DIR = Path("/root/reference")


def ref_lib():
    """oracle/_ref/libref_ptam.so — the reference's OWN sources (Bundle.cc, ATANCamera.cc, ...) compiled
    where they lie against the header stand-ins in oracle/shim/ (Makefile.ref).  Built when
    /root/reference is present; on the GPU box only the prebuilt file is used.  None when neither."""
    so = ORACLE_DIR / "_ref" / "libref_ptam.so"
    if REFERENCE_DIR.exists():
        r = subprocess.run(["make", "-C", str(ORACLE_DIR), "-f", "Makefile.ref"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building oracle/_ref failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    if not so.exists():
        return None
    if "ref" not in _lib:
        _lib["ref"] = Lib(so, "ref_")
    return _lib["ref"]
