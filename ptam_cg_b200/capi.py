"""ctypes binding of the C-ABI in include/ptam_b200.h (plumbing for tests and bench.py).

The binding is generic over (shared library, symbol prefix): the product is
``libptam_b200.so`` / ``ptam_``; the CPU oracle (test infrastructure under ``oracle/``) exports the
same signatures with prefix ``orc_`` and is bound by ``oracle/binding.py`` — never from here.
The product path has no CPU fallback: if the CUDA library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

LEVELS = 4
PT_IN_IMAGE, PT_IN_PVS, PT_SEARCHED, PT_FOUND, PT_SUBPIX, PT_TEMPLATE_BAD = 1, 2, 4, 8, 16, 32

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "csrc" / "libptam_b200.so"


class TrackerParams(C.Structure):
    _fields_ = [
        ("coarse_min", C.c_int32), ("coarse_max", C.c_int32), ("coarse_range", C.c_int32),
        ("coarse_subpix_its", C.c_int32), ("disable_coarse", C.c_int32),
        ("max_patches_per_frame", C.c_int32), ("mestimator", C.c_int32),
        ("use_constant_velocity", C.c_int32), ("coarse_min_velocity", C.c_double),
        ("quality_good", C.c_double), ("quality_lost", C.c_double),
        ("use_rotation_estimator", C.c_int32), ("reserved0", C.c_int32), ("rotation_estimator_blur", C.c_double),
    ]


class TrackerState(C.Structure):
    _fields_ = [
        ("se3_cam_from_world", C.c_double * 12), ("velocity", C.c_double * 6),
        ("msd_scaled_velocity_magnitude", C.c_double), ("scene_depth_mean", C.c_double),
        ("scene_depth_sigma", C.c_double), ("just_recovered_so_use_coarse", C.c_int32),
        ("tracking_quality", C.c_int32), ("lost_frames", C.c_int32), ("frame", C.c_int32),
    ]


class TrackResult(C.Structure):
    _fields_ = [
        ("se3_cam_from_world", C.c_double * 12), ("scene_depth_mean", C.c_double),
        ("scene_depth_sigma", C.c_double), ("meas_attempted", C.c_int32 * 4),
        ("meas_found", C.c_int32 * 4), ("n_corners", C.c_int32 * 4), ("did_coarse", C.c_int32),
        ("n_coarse", C.c_int32), ("n_level3", C.c_int32), ("n_fine", C.c_int32),
        ("tracking_quality", C.c_int32), ("quality_needs_kf_distance", C.c_int32),
        ("n_pvs", C.c_int32 * 4), ("n_candidates", C.c_int32),
        ("recovery", C.c_int32), ("reloc_keyframe", C.c_int32), ("reserved1", C.c_int32), ("reloc_score", C.c_double),
    ]


class BundleParams(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int32), ("mestimator", C.c_int32),
        ("update_squared_convergence", C.c_double), ("min_tukey_sigma", C.c_double),
    ]


class BundleStats(C.Structure):
    _fields_ = [
        ("accepted", C.c_int32), ("lambda_trials", C.c_int32), ("lm_steps", C.c_int32),
        ("converged", C.c_int32), ("hit_max_iterations", C.c_int32), ("n_outliers", C.c_int32),
        ("sigma_squared", C.c_double), ("lambda_", C.c_double), ("last_error", C.c_double),
        ("last_new_error", C.c_double),
    ]


TRACKER_SYMBOLS = [
    "tracker_default_params", "tracker_create", "tracker_destroy", "tracker_last_error",
    "tracker_add_keyframe", "tracker_set_map", "tracker_set_state", "tracker_get_state",
    "tracker_make_keyframes", "tracker_track_frames", "tracker_synchronize", "tracker_get_level",
    "tracker_level_size", "tracker_get_points", "tracker_get_templates", "tracker_get_sbi",
    "tracker_keyframe_rest", "tracker_get_level_rest", "tracker_refind_in_keyframes",
    "tracker_get_iteration_set", "tracker_epipolar_search", "tracker_set_keyframe_pose",
    "patch_search_batch", "patch_get_results", "pose_update",
]
BUNDLE_SYMBOLS = [
    "bundle_default_params", "bundle_create", "bundle_destroy", "bundle_last_error",
    "bundle_add_camera", "bundle_add_point", "bundle_add_meas", "bundle_add_cameras",
    "bundle_add_points", "bundle_add_measurements", "bundle_set_shard", "bundle_compute",
    "bundle_begin", "bundle_lm_step", "bundle_converged", "bundle_get_point", "bundle_get_camera",
    "bundle_get_points", "bundle_get_cameras", "bundle_get_outliers", "bundle_get_stats",
    "bundle_get_reduced_system", "bundle_synchronize",
    "bundle_recompute", "bundle_update_camera", "bundle_update_point",
]
# exported by the product only (CUDA plumbing)
PRODUCT_ONLY_SYMBOLS = [
    "global_last_error", "tracker_track_frames_device", "tracker_submit_frames", "tracker_submit_frames_device", "tracker_collect", "tracker_cuda_stream",
    "tracker_launch_count", "tracker_set_profiling", "tracker_get_kernel_times",
    "bundle_cuda_stream", "bundle_launch_count", "bundle_solve_schedule",
    "nccl_unique_id", "nccl_comm_create", "nccl_comm_destroy", "bundle_init_shard", "bundle_shard_plan",
    "bundle_set_profiling", "bundle_get_phase_times",
]
BUNDLE_PHASES = ["project", "select", "jacobian", "vinv_init", "schur", "allreduce", "solve", "update_newerror", "reserved",
                 "commit_erase", "begin_setup", "host_control_and_sync"]
TRACKER_KERNELS = ["k_fast2_l0", "k_fast2_l123", "k_compact", "k_sbi+k_pvs_select", "k_search_coarse", "k_pose_coarse",
                   "k_search_fine", "k_pose_fine"]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _bp(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Lib:
    """A loaded library exposing <prefix>tracker_* / <prefix>bundle_*."""

    def __init__(self, path, prefix):
        path = str(path)
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                "There is no CPU fallback.")
        self.path, self.prefix = path, prefix
        self.cdll = C.CDLL(path)
        self._declare()

    def fn(self, name):
        return getattr(self.cdll, self.prefix + name)

    def has(self, name):
        return hasattr(self.cdll, self.prefix + name)

    def _declare(self):
        vp, i, d = C.c_void_p, C.c_int, C.c_double
        P = C.POINTER
        sig = {
            "tracker_default_params": (None, [P(TrackerParams)]),
            "tracker_create": (vp, [i, P(d), i, i, i, P(TrackerParams)]),
            "tracker_destroy": (None, [vp]),
            "tracker_last_error": (C.c_char_p, [vp]),
            "tracker_add_keyframe": (i, [vp, P(C.c_uint8), i]),
            "tracker_set_map": (i, [vp, i, i, P(d), P(d), P(d), P(C.c_int32), P(C.c_int32), P(C.c_int32)]),
            "tracker_set_state": (i, [vp, i, P(TrackerState)]),
            "tracker_set_keyframe_pose": (i, [vp, i, P(d)]),
            "tracker_get_state": (i, [vp, i, P(TrackerState)]),
            "tracker_make_keyframes": (i, [vp, P(vp), i]),
            "tracker_track_frames": (i, [vp, P(vp), i, P(TrackResult)]),
            "tracker_track_frames_device": (i, [vp, vp, C.c_size_t, i, P(TrackResult)]),
            "tracker_submit_frames": (i, [vp, P(vp), i]),
            "tracker_submit_frames_device": (i, [vp, vp, C.c_size_t, i, vp]),
            "tracker_collect": (i, [vp, P(TrackResult)]),
            "tracker_synchronize": (i, [vp]),
            "tracker_cuda_stream": (vp, [vp]),
            "tracker_launch_count": (C.c_int64, [vp]),
            "tracker_set_profiling": (i, [vp, i]),
            "tracker_get_kernel_times": (i, [vp, P(d), P(C.c_int64)]),
            "tracker_get_level": (i, [vp, i, i, P(C.c_uint8), P(C.c_int32), i, P(C.c_int32)]),
            "tracker_level_size": (i, [vp, i, P(i), P(i)]),
            "tracker_get_points": (i, [vp, i, P(C.c_int32), P(C.c_int32), P(d), P(d), P(C.c_int32), P(C.c_int32)]),
            "tracker_get_templates": (i, [vp, i, P(C.c_uint8), P(C.c_int32)]),
            "tracker_get_sbi": (i, [vp, i, P(C.c_float), i, P(d), P(d)]),
            "tracker_keyframe_rest": (i, [vp, i, d]),
            "tracker_refind_in_keyframes": (i, [vp, P(vp), i, P(d)]),
            "tracker_get_level_rest": (i, [vp, i, i, P(C.c_int32), i, P(C.c_int32), P(d), i, P(i)]),
            "tracker_get_iteration_set": (i, [vp, i, P(C.c_int32), i]),
            "tracker_epipolar_search": (i, [vp, i, i, i, P(d), d, d, P(d), d, i, P(C.c_int32), P(C.c_int32), P(C.c_int32), P(d)]),
            "patch_search_batch": (i, [vp, P(d), C.c_uint, i]),
            "patch_get_results": (i, [vp, i, P(C.c_int32), P(d), P(C.c_int32), P(C.c_int32), P(d), P(C.c_int32)]),
            "pose_update": (i, [vp, d, i, P(d), P(C.c_int32)]),
            "global_last_error": (C.c_char_p, []),
            "bundle_default_params": (None, [P(BundleParams)]),
            "bundle_create": (vp, [i, P(d), i, i, P(BundleParams)]),
            "bundle_destroy": (None, [vp]),
            "bundle_last_error": (C.c_char_p, [vp]),
            "bundle_add_camera": (i, [vp, P(d), i]),
            "bundle_add_point": (i, [vp, P(d)]),
            "bundle_add_meas": (i, [vp, i, i, P(d), d]),
            "bundle_add_cameras": (i, [vp, i, P(d), P(C.c_int32)]),
            "bundle_add_points": (i, [vp, i, P(d)]),
            "bundle_add_measurements": (i, [vp, i, P(C.c_int32), P(C.c_int32), P(d), P(d)]),
            "bundle_set_shard": (i, [vp, i, i, vp]),
            "nccl_unique_id": (i, [P(C.c_ubyte)]),
            "nccl_comm_create": (vp, [i, i, i, P(C.c_ubyte)]),
            "nccl_comm_destroy": (None, [vp]),
            "bundle_init_shard": (i, [vp, i, i, P(C.c_ubyte)]),
            "bundle_shard_plan": (i, [i, i, P(C.c_int32), i, P(C.c_int32)]),
            "bundle_compute": (i, [vp, P(C.c_ubyte)]),
            "bundle_begin": (i, [vp]),
            "bundle_recompute": (i, [vp, P(C.c_ubyte)]),
            "bundle_update_camera": (i, [vp, i, P(d)]),
            "bundle_update_point": (i, [vp, i, P(d)]),
            "bundle_lm_step": (i, [vp, P(C.c_ubyte)]),
            "bundle_converged": (i, [vp]),
            "bundle_get_point": (i, [vp, i, P(d)]),
            "bundle_get_camera": (i, [vp, i, P(d)]),
            "bundle_get_points": (i, [vp, P(d)]),
            "bundle_get_cameras": (i, [vp, P(d)]),
            "bundle_get_outliers": (i, [vp, P(C.c_int32), i]),
            "bundle_get_stats": (i, [vp, P(BundleStats)]),
            "bundle_get_reduced_system": (i, [vp, P(d), P(d), i]),
            "bundle_synchronize": (i, [vp]),
            "bundle_cuda_stream": (vp, [vp]),
            "bundle_launch_count": (C.c_int64, [vp]),
            "bundle_set_profiling": (i, [vp, i]),
            "bundle_get_phase_times": (i, [vp, P(d), P(C.c_int64)]),
        }
        for name, (res, args) in sig.items():
            if self.has(name):
                f = self.fn(name)
                f.restype, f.argtypes = res, args


_product = None


def product_lib() -> Lib:
    """The CUDA product library. Raises if it has not been built — never falls back."""
    global _product
    if _product is None:
        _product = Lib(os.environ.get("PTAM_B200_LIB", LIB_PATH), "ptam_")  # the override is for kernel experiments (scripts/lab)
    return _product


NCCL_UNIQUE_ID_BYTES = 128


def nccl_unique_id(lib: Lib) -> bytes:
    buf = (C.c_ubyte * NCCL_UNIQUE_ID_BYTES)()
    if lib.fn("nccl_unique_id")(buf) != 0:
        raise PtamError(lib.fn("global_last_error")().decode())
    return bytes(buf)


def nccl_comm_create(lib: Lib, device, rank, world, unique_id: bytes):
    """One communicator per process, shared by every sharded Bundle (pass it to Bundle.set_shard)."""
    buf = (C.c_ubyte * NCCL_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
    comm = lib.fn("nccl_comm_create")(int(device), int(rank), int(world), buf)
    if not comm:
        raise PtamError(lib.fn("global_last_error")().decode())
    return comm


def shard_plan(lib: Lib, n_points, meas_point, world):
    """point_begin[world + 1]: shard r owns points [point_begin[r], point_begin[r + 1])."""
    mp = _i32(meas_point)
    out = np.zeros(world + 1, np.int32)
    if lib.fn("bundle_shard_plan")(int(n_points), len(mp), _ip(mp), int(world), _ip(out)) != 0:
        raise PtamError("bad shard plan arguments")
    return out


CAMERA_PARAMS = np.array([1.0803, 1.43987, 0.519983, 0.548655, 0.244943])  # config/camera.cfg:7


class PtamError(RuntimeError):
    pass


class Tracker:
    """Batch of n_streams trackers behind ptam_tracker_* (mirrors Tracker::TrackFrame,
    reference src/Tracker.cc:86-188)."""

    def __init__(self, lib: Lib, width, height, n_streams=1, cam_params=CAMERA_PARAMS, device=0, **params):
        self.lib, self.W, self.H, self.S = lib, int(width), int(height), int(n_streams)
        p = TrackerParams()
        lib.fn("tracker_default_params")(C.byref(p))
        for k, v in params.items():
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, v)
        self.params = p
        cp = _f64(cam_params)
        self.h = lib.fn("tracker_create")(device, _dp(cp), self.W, self.H, self.S, C.byref(p))
        if not self.h:
            msg = lib.fn("global_last_error")().decode() if lib.has("global_last_error") else "create failed"
            raise PtamError(msg)
        self.n_points = [0] * self.S

    def close(self):
        if self.h:
            self.lib.fn("tracker_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise PtamError(f"rc={rc}: {self.lib.fn('tracker_last_error')(self.h).decode()}")
        return rc

    def add_keyframe(self, image):
        image = np.ascontiguousarray(image, dtype=np.uint8)
        assert image.shape == (self.H, self.W)
        return self._chk(self.lib.fn("tracker_add_keyframe")(self.h, _bp(image), self.W))

    def set_keyframe_pose(self, kf, pose12):
        """KeyFrame::se3CfromW of a stored keyframe; once every keyframe has one the relocaliser branch is on."""
        self._chk(self.lib.fn("tracker_set_keyframe_pose")(self.h, int(kf), _dp(_f64(pose12).reshape(12))))

    def set_map(self, stream, m):
        w, r, dn = _f64(m["world_pos"]), _f64(m["pixel_right_w"]), _f64(m["pixel_down_w"])
        k, l, c = _i32(m["src_kf"]), _i32(m["src_level"]), _i32(m["ir_center"])
        n = len(k)
        self._chk(self.lib.fn("tracker_set_map")(self.h, stream, n, _dp(w), _dp(r), _dp(dn), _ip(k), _ip(l), _ip(c)))
        self.n_points[stream] = n

    def set_state(self, stream, pose12=None, velocity=None, msd=None, depth_mean=None, just_recovered=None, state=None):
        st = state if state is not None else self.get_state(stream)
        if pose12 is not None:
            st.se3_cam_from_world[:] = list(_f64(pose12).reshape(12))
        if velocity is not None:
            st.velocity[:] = list(_f64(velocity).reshape(6))
        if msd is not None:
            st.msd_scaled_velocity_magnitude = msd
        if depth_mean is not None:
            st.scene_depth_mean = depth_mean
        if just_recovered is not None:
            st.just_recovered_so_use_coarse = int(just_recovered)
        self._chk(self.lib.fn("tracker_set_state")(self.h, stream, C.byref(st)))

    def get_state(self, stream) -> TrackerState:
        st = TrackerState()
        self._chk(self.lib.fn("tracker_get_state")(self.h, stream, C.byref(st)))
        return st

    def _image_ptrs(self, images):
        imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        assert len(imgs) == self.S and all(im.shape == (self.H, self.W) for im in imgs)
        arr = (C.c_void_p * self.S)(*[im.ctypes.data for im in imgs])
        return imgs, arr

    def make_keyframes(self, images):
        imgs, arr = self._image_ptrs(images)
        self._chk(self.lib.fn("tracker_make_keyframes")(self.h, arr, self.W))

    def track_frames(self, images):
        imgs, arr = self._image_ptrs(images)
        res = (TrackResult * self.S)()
        self._chk(self.lib.fn("tracker_track_frames")(self.h, arr, self.W, res))
        return list(res)

    def track_frames_device(self, dptr, frame_pitch, stride, want_results=False):
        res = (TrackResult * self.S)() if want_results else None
        self._chk(self.lib.fn("tracker_track_frames_device")(self.h, C.c_void_p(dptr), frame_pitch, stride, res))
        return list(res) if want_results else None

    def synchronize(self):
        self._chk(self.lib.fn("tracker_synchronize")(self.h))

    def cuda_stream(self):
        return self.lib.fn("tracker_cuda_stream")(self.h)

    def launch_count(self):
        return int(self.lib.fn("tracker_launch_count")(self.h))

    def set_profiling(self, on):
        self._chk(self.lib.fn("tracker_set_profiling")(self.h, int(on)))

    def kernel_times(self):
        """{kernel name: (total ms, launches)} accumulated since set_profiling(True)."""
        ms = (C.c_double * 8)()
        n = (C.c_int64 * 8)()
        self._chk(self.lib.fn("tracker_get_kernel_times")(self.h, ms, n))
        return {k: (ms[j], n[j]) for j, k in enumerate(TRACKER_KERNELS)}

    def track_frames_ptrs(self, ptrs, stride, want_results=True):
        """track_frames on raw host pointers (e.g. pinned torch tensors), one per stream."""
        arr = (C.c_void_p * self.S)(*ptrs)
        res = (TrackResult * self.S)() if want_results else None
        self._chk(self.lib.fn("tracker_track_frames")(self.h, arr, stride, res))
        return list(res) if want_results else None

    def submit_ptrs(self, ptrs, stride):
        """Pipelined track_frames: enqueue H2D + kernels for one batch of raw host pointers."""
        arr = (C.c_void_p * self.S)(*ptrs)
        self._chk(self.lib.fn("tracker_submit_frames")(self.h, arr, stride))

    def ptr_array(self, ptrs):
        """ctypes pointer array for submit_array (build once per batch, outside any timed loop)."""
        return (C.c_void_p * self.S)(*ptrs)

    def submit_device(self, dptr, frame_pitch, stride, ready_event=None):
        """Pipelined track_frames_device: frames resident in device memory, complete now (or at ready_event)."""
        self._chk(self.lib.fn("tracker_submit_frames_device")(self.h, C.c_void_p(dptr), frame_pitch, stride,
                                                              C.c_void_p(ready_event) if ready_event else None))

    def submit_array(self, arr, stride):
        self._chk(self.lib.fn("tracker_submit_frames")(self.h, arr, stride))

    def result_buffer(self):
        return (TrackResult * self.S)()

    def collect_into(self, res):
        self._chk(self.lib.fn("tracker_collect")(self.h, res))
        return res

    def collect(self, want_results=True):
        res = (TrackResult * self.S)() if want_results else None
        self._chk(self.lib.fn("tracker_collect")(self.h, res))
        return list(res) if want_results else None

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        self.lib.fn("tracker_level_size")(self.h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def get_level(self, stream, level):
        w, h = self.level_size(level)
        pix = np.zeros((h, w), np.uint8)
        lut = np.zeros(h, np.int32)
        n = self._chk(self.lib.fn("tracker_get_level")(self.h, stream, level, _bp(pix), None, 0, _ip(lut)))
        xy = np.zeros((max(n, 1), 2), np.int32)
        self._chk(self.lib.fn("tracker_get_level")(self.h, stream, level, None, _ip(xy), n, None))
        return pix, xy[:n], lut

    def get_points(self, stream):
        n = self.n_points[stream]
        flags, level = np.zeros(n, np.int32), np.zeros(n, np.int32)
        found, image = np.zeros((n, 2)), np.zeros((n, 2))
        outl, inl = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._chk(self.lib.fn("tracker_get_points")(self.h, stream, _ip(flags), _ip(level), _dp(found), _dp(image), _ip(outl), _ip(inl)))
        return dict(flags=flags, level=level, v2_found=found, v2_image=image, outliers=outl, inliers=inl)

    def get_templates(self, stream):
        n = self.n_points[stream]
        t, s = np.zeros((n, 64), np.uint8), np.zeros((n, 2), np.int32)
        self._chk(self.lib.fn("tracker_get_templates")(self.h, stream, _bp(t), _ip(s)))
        return t, s

    def refind_in_keyframes(self, images, poses12):
        """MapMaker::ReFindInSingleKeyFrame, one keyframe (image, se3CfromW) per stream; results via get_points."""
        imgs, arr = self._image_ptrs(images)
        p = _f64(poses12).reshape(self.S, 12)
        self._chk(self.lib.fn("tracker_refind_in_keyframes")(self.h, arr, self.W, _dp(p)))

    def patch_search_batch(self, poses12, search_range, subpix_its):
        """PatchFinder steps 1-5 (PatchFinder.h:54-98) for every map point of every stream against the stream's
        current frame at the given poses; results via patch_results / get_templates."""
        p = _f64(poses12).reshape(self.S, 12)
        self._chk(self.lib.fn("patch_search_batch")(self.h, _dp(p), int(search_range), int(subpix_its)))

    def patch_results(self, stream):
        n = self._chk(self.lib.fn("patch_get_results")(self.h, stream, None, None, None, None, None, None))
        m = max(n, 1)
        out = dict(level=np.zeros(m, np.int32), warp_inverse=np.zeros((m, 4)), template_bad=np.zeros(m, np.int32),
                   found=np.zeros(m, np.int32), pos=np.zeros((m, 2)), subpix=np.zeros(m, np.int32))
        self._chk(self.lib.fn("patch_get_results")(self.h, stream, _ip(out["level"]), _dp(out["warp_inverse"]), _ip(out["template_bad"]),
                                                   _ip(out["found"]), _dp(out["pos"]), _ip(out["subpix"])))
        return {k: v[:n] for k, v in out.items()}

    def pose_update(self, override_sigma_squared=0.0, mark_outliers=False):
        """Tracker::CalcPoseUpdate once per stream over the points found by the last patch_search_batch:
        (mu (S, 6), n_found (S,))."""
        mu, nf = np.zeros((self.S, 6)), np.zeros(self.S, np.int32)
        self._chk(self.lib.fn("pose_update")(self.h, float(override_sigma_squared), int(bool(mark_outliers)), _dp(mu), _ip(nf)))
        return mu, nf

    def epipolar_search(self, stream, level, src_kf, src_pose12, src_depth_mean, src_depth_sigma, target_pose12, wiggle_scale, cand_xy):
        """MapMaker::AddPointEpipolar up to the sub-pixel target position for every candidate (irLevelPos in the
        stored source keyframe's level); target = the stream's current frame.  Returns (found (n,), best corner
        index (n,), sub-pixel level-zero position in the target (n, 2))."""
        c = _i32(cand_xy).reshape(-1, 2)
        n = len(c)
        found, best, sub = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32), np.zeros((max(n, 1), 2))
        self._chk(self.lib.fn("tracker_epipolar_search")(self.h, stream, level, src_kf, _dp(_f64(src_pose12).reshape(12)),
                                                        float(src_depth_mean), float(src_depth_sigma),
                                                        _dp(_f64(target_pose12).reshape(12)), float(wiggle_scale), n, _ip(c),
                                                        _ip(found), _ip(best), _dp(sub)))
        return found[:n], best[:n], sub[:n]

    def keyframe_rest(self, stream, min_shi_tomasi_score=70.0):
        """KeyFrame::MakeKeyFrame_Rest for the stream's current frame; returns per level
        (vMaxCorners (n,2) int32, candidate positions (m,2) int32, candidate Shi-Tomasi scores (m,))."""
        self._chk(self.lib.fn("tracker_keyframe_rest")(self.h, stream, float(min_shi_tomasi_score)))
        out = []
        for l in range(LEVELS):
            nc = C.c_int()
            nm = self._chk(self.lib.fn("tracker_get_level_rest")(self.h, stream, l, None, 0, None, None, 0, C.byref(nc)))
            mx, cx, cs = np.zeros((max(nm, 1), 2), np.int32), np.zeros((max(nc.value, 1), 2), np.int32), np.zeros(max(nc.value, 1))
            self._chk(self.lib.fn("tracker_get_level_rest")(self.h, stream, l, _ip(mx), nm, _ip(cx), _dp(cs), nc.value, C.byref(nc)))
            out.append((mx[:nm], cx[:nc.value], cs[:nc.value]))
        return out

    def get_sbi(self, stream):
        """(mimTemplate as (h, w) float32, so3 rotation estimate (3,), final ESM score) of the last frame."""
        n = self.lib.fn("tracker_get_sbi")(self.h, stream, None, 0, None, None)
        tmpl, rot, score = np.zeros(max(n, 1), np.float32), np.zeros(3), C.c_double()
        self._chk(self.lib.fn("tracker_get_sbi")(self.h, stream, tmpl.ctypes.data_as(C.POINTER(C.c_float)), n, _dp(rot), C.byref(score)))
        w3, h3 = self.level_size(3)
        return tmpl[:n].reshape(h3 // 2, w3 // 2), rot, score.value

    def get_iteration_set(self, stream):
        n = self.n_points[stream]
        idx = np.zeros(max(n, 1), np.int32)
        k = self._chk(self.lib.fn("tracker_get_iteration_set")(self.h, stream, _ip(idx), n))
        return idx[:k]


class Bundle:
    """Mirror of class Bundle (reference include/Bundle.h:105-156) over ptam_bundle_*."""

    def __init__(self, lib: Lib, width=640, height=480, cam_params=CAMERA_PARAMS, device=0, **params):
        self.lib = lib
        p = BundleParams()
        lib.fn("bundle_default_params")(C.byref(p))
        for k, v in params.items():
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, v)
        cp = _f64(cam_params)
        self.h = lib.fn("bundle_create")(device, _dp(cp), int(width), int(height), C.byref(p))
        if not self.h:
            msg = lib.fn("global_last_error")().decode() if lib.has("global_last_error") else "create failed"
            raise PtamError(msg)
        self.n_cams = self.n_points = 0

    def close(self):
        if self.h:
            self.lib.fn("bundle_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise PtamError(f"rc={rc}: {self.lib.fn('bundle_last_error')(self.h).decode()}")
        return rc

    # reference names: AddCamera / AddPoint / AddMeas / Compute / Converged / GetPoint / GetCamera
    def AddCamera(self, se3, fixed):
        a = _f64(se3).reshape(12)
        self.n_cams += 1
        return self._chk(self.lib.fn("bundle_add_camera")(self.h, _dp(a), int(bool(fixed))))

    def AddPoint(self, xyz):
        a = _f64(xyz).reshape(3)
        self.n_points += 1
        return self._chk(self.lib.fn("bundle_add_point")(self.h, _dp(a)))

    def AddMeas(self, cam, point, uv, sigma_squared):
        a = _f64(uv).reshape(2)
        self._chk(self.lib.fn("bundle_add_meas")(self.h, int(cam), int(point), _dp(a), float(sigma_squared)))

    def add_graph(self, g):
        se3, fixed = _f64(g["cam_se3"]), _i32(g["cam_fixed"])
        self._chk(self.lib.fn("bundle_add_cameras")(self.h, len(fixed), _dp(se3), _ip(fixed)))
        pts = _f64(g["points"])
        self._chk(self.lib.fn("bundle_add_points")(self.h, len(pts), _dp(pts)))
        mc, mp, uv, s2 = _i32(g["meas_cam"]), _i32(g["meas_point"]), _f64(g["meas_uv"]), _f64(g["meas_sigma_sq"])
        self._chk(self.lib.fn("bundle_add_measurements")(self.h, len(mc), _ip(mc), _ip(mp), _dp(uv), _dp(s2)))
        self.n_cams += len(fixed)
        self.n_points += len(pts)

    def set_shard(self, rank, world, comm):
        self._chk(self.lib.fn("bundle_set_shard")(self.h, rank, world, C.c_void_p(comm)))

    def init_shard(self, rank, world, unique_id: bytes):
        """Join the NCCL communicator described by `unique_id` (from nccl_unique_id() on rank 0)."""
        buf = (C.c_ubyte * NCCL_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self._chk(self.lib.fn("bundle_init_shard")(self.h, rank, world, buf))

    def Compute(self, abort=None):
        return self._chk(self.lib.fn("bundle_compute")(self.h, abort))

    def Recompute(self, abort=None):
        """Bundle::Compute again on the resident graph (adjusted state, outliers erased): SURVEY 8f rank 4."""
        return self._chk(self.lib.fn("bundle_recompute")(self.h, abort))

    def update_camera(self, n, se3):
        self._chk(self.lib.fn("bundle_update_camera")(self.h, int(n), _dp(_f64(se3).reshape(12))))

    def update_point(self, n, xyz):
        self._chk(self.lib.fn("bundle_update_point")(self.h, int(n), _dp(_f64(xyz).reshape(3))))

    def begin(self):
        self._chk(self.lib.fn("bundle_begin")(self.h))

    def lm_step(self, abort=None):
        return self._chk(self.lib.fn("bundle_lm_step")(self.h, abort))

    def Converged(self):
        return bool(self.lib.fn("bundle_converged")(self.h))

    def GetPoint(self, n):
        a = np.zeros(3)
        self._chk(self.lib.fn("bundle_get_point")(self.h, n, _dp(a)))
        return a

    def GetCamera(self, n):
        a = np.zeros(12)
        self._chk(self.lib.fn("bundle_get_camera")(self.h, n, _dp(a)))
        return a

    def get_points(self):
        a = np.zeros((self.n_points, 3))
        self._chk(self.lib.fn("bundle_get_points")(self.h, _dp(a)))
        return a

    def get_cameras(self):
        a = np.zeros((self.n_cams, 12))
        self._chk(self.lib.fn("bundle_get_cameras")(self.h, _dp(a)))
        return a

    def GetOutlierMeasurements(self):
        n = self._chk(self.lib.fn("bundle_get_outliers")(self.h, None, 0))
        a = np.zeros((max(n, 1), 2), np.int32)
        self._chk(self.lib.fn("bundle_get_outliers")(self.h, _ip(a), n))
        return a[:n]

    def stats(self) -> BundleStats:
        s = BundleStats()
        self._chk(self.lib.fn("bundle_get_stats")(self.h, C.byref(s)))
        return s

    def reduced_system(self, n):
        S, vE = np.zeros((n, n)), np.zeros(n)
        k = self._chk(self.lib.fn("bundle_get_reduced_system")(self.h, _dp(S), _dp(vE), n))
        return S[:k, :k], vE[:k]

    def synchronize(self):
        self._chk(self.lib.fn("bundle_synchronize")(self.h))

    def launch_count(self):
        return int(self.lib.fn("bundle_launch_count")(self.h))

    def cuda_stream(self):
        return self.lib.fn("bundle_cuda_stream")(self.h)

    def set_profiling(self, on):
        self._chk(self.lib.fn("bundle_set_profiling")(self.h, int(on)))

    def phase_times(self):
        """{phase: (total ms, count)} accumulated since set_profiling(True)."""
        ms, n = (C.c_double * len(BUNDLE_PHASES))(), (C.c_int64 * len(BUNDLE_PHASES))()
        self._chk(self.lib.fn("bundle_get_phase_times")(self.h, ms, n))
        return {k: (ms[j], n[j]) for j, k in enumerate(BUNDLE_PHASES)}
